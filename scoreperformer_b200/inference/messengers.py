"""Note tuples -> timed MIDI messages (the step after rendering; reference inference/messengers.py:20-363).

Same classes, arguments and return values as the reference, on top of any tokenizer-like object that offers `vocab_types_idx`,
`zero_token`, `decode_token_type`, `compute_ticks` and `config` (`token_tables.TokenTables` or the reference's own tokenizer).
Host numpy throughout: a window of a rendering is a few dozen notes, and the tempo tracking of SPMuple2 is a recurrence over onsets.
What differs from the reference is the shape of the code, not the arithmetic: notes are grouped by onset once (a stable sort) instead
of one boolean mask per onset, and both messengers share the message packing; every floating-point expression keeps the reference's
operand order so that times agree bit for bit (tests/test_inference_host.py).

A message row is `(time_or_tick, 144, pitch, velocity)`; note-offs carry velocity 0.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import native
from .token_tables import DEFAULT_TEMPO, NOTE_ON_MIDI_EVENT, TokenTables


@dataclass
class IntermediateData:
    tempos: Optional[np.ndarray] = None                  # rows (tempo, tick, time) of the tempo map so far


@dataclass
class SPMuple2IntermediateData(IntermediateData):
    initial_tempo: float = DEFAULT_TEMPO
    onset_pairs: Optional[np.ndarray] = None             # rows (score tick, performed time, notes averaged) per onset


def _drop_repeats(tempos: np.ndarray) -> np.ndarray:
    """Keep the last row of every run of equal ticks, then the first row of every run of equal tempi (messengers.py:139-143)."""
    ticks = np.concatenate([tempos[:, 1], [-1]])
    tempos = tempos[(ticks[1:] - ticks[:-1]) != 0]
    bpm = np.concatenate([[-1], tempos[:, 0]])
    return tempos[(bpm[1:] - bpm[:-1]) != 0]


class SPMupleMessenger:
    """Tick-based decoding: score ticks shifted by the onset deviation, a tempo map tied to beat (or bar) starts."""

    def __init__(self, tokenizer):
        self.tokenizer = tokenizer
        self.beat_resolution = max(tokenizer.config.beat_res.values())

    # ------------------------------------------------------------ shared pieces
    def _pack(self, tokens, on, off, note_attributes, note_on_events, note_off_events):
        assert note_on_events or note_off_events
        parts = []
        if note_attributes:
            pitch = self.tokenizer.decode_token_type(tokens, "Pitch")
            velocity = self.tokenizer.decode_token_type(tokens, "Velocity")
            kind = np.full_like(pitch, NOTE_ON_MIDI_EVENT)
            if note_on_events:
                parts.append(np.stack([on, kind, pitch, velocity], axis=-1))
            if note_off_events:
                parts.append(np.stack([off, kind, pitch, np.zeros(velocity.shape[0])], axis=-1))
        else:
            parts = [x for x, keep in ((on, note_on_events), (off, note_off_events)) if keep]
        return np.concatenate(parts, axis=0)

    def _performed_ticks(self, tokens, on, durations):
        """(onset, offset) ticks after the performance fields are applied (messengers.py:46-72)."""
        tok, p = self.tokenizer, self.tokenizer.config.additional_params
        fields = tok.vocab_types_idx
        if "RelOnsetDev" not in fields and "OnsetDev" not in fields:       # score-only vocabulary: nothing to apply
            return on, on + durations
        shifts = tok.decode_token_type(tokens, "PositionShift") if p["use_position_shifts"] else tok.compute_position_shifts(on)
        if p["rel_onset_dev"]:
            shifts[shifts == 0] = 1
            dev = tok.decode_token_type(tokens, "RelOnsetDev") * shifts
        else:
            dev = tok.decode_token_type(tokens, "OnsetDev")
        on = np.maximum(0, on + dev)
        if p["rel_perf_duration"]:
            held = tok.decode_token_type(tokens, "RelPerfDuration") * durations
        else:
            held = tok.decode_token_type(tokens, "PerfDuration")
        return on, on + held

    # ------------------------------------------------------------ API
    def tokens_to_messages(self, tokens: np.ndarray, note_attributes: bool = True, note_on_events: bool = True,
                           note_off_events: bool = True, intermediates: Optional[IntermediateData] = None,
                           return_intermediates: bool = False, to_times: bool = True, sort: bool = True):
        tok = self.tokenizer
        grid = tok.compute_ticks(tokens, self.beat_resolution, compute_beat_ticks=True)
        durations = tok.decode_token_type(tokens, "Duration")
        on, off = self._performed_ticks(tokens, grid["note_on"].astype(float), durations)

        # tempo map of this chunk, continued from the previous chunk's last row
        col = tokens[:, tok.vocab_types_idx["Tempo"]]
        change = np.concatenate([[0], np.flatnonzero(col[1:] != col[:-1]) + 1])
        bpm = tok.decode_token_type(tokens[change], "Tempo")
        before = intermediates.tempos if intermediates is not None else None
        resumes_changed = before is not None and before[-1, 0] != bpm[0]
        if resumes_changed:
            bpm = np.concatenate([[before[-1, 0]], bpm])
        tick0, time0 = (0, 0.) if before is None else (before[-1, 1], before[-1, 2])

        marks = grid["bar"] if tok.config.additional_params["bar_tempos"] else grid["beat"]
        snap = lambda t: marks[np.minimum(np.searchsorted(marks, t), marks.shape[0] - 1)]     # next beat / bar start
        at = snap(on[change])
        at[0] = tick0
        if resumes_changed:
            at = np.concatenate([[at[0]], [snap(on[0])], at[1:]])
        times = np.cumsum(np.concatenate([[time0], np.diff(at) / self.beat_resolution * 60 / bpm[:-1]]))
        tempo_map = np.stack([bpm, at, times], axis=-1)

        messages = self._pack(tokens, on, off, note_attributes, note_on_events, note_off_events)
        if to_times:
            messages = self.messages_to_times(messages, tempo_map, sort=sort)
        elif sort:
            messages = self.sort_messages(messages)
        if not return_intermediates:
            return messages
        merged = tempo_map if before is None else np.concatenate([before, tempo_map[1:]], axis=0)
        return messages, IntermediateData(tempos=_drop_repeats(merged))

    def messages_to_times(self, messages: np.ndarray, tempos: np.ndarray, sort: bool = True, inplace: bool = True):
        """Ticks -> seconds through the piecewise-constant tempo map (messengers.py:149-173)."""
        flat = messages.ndim == 1
        ticks = messages if flat else messages[:, 0]
        seg = np.searchsorted(tempos[:, 1], ticks, side="right") - 1
        bpm, at, t0 = tempos[seg, 0], tempos[seg, 1], tempos[seg, 2]
        seconds = t0 + (ticks - at) / self.beat_resolution * 60 / bpm
        if not inplace:
            messages = messages.copy()
        if flat:
            messages[:] = seconds
        else:
            messages[:, 0] = seconds
        return self.sort_messages(messages) if sort else messages

    @staticmethod
    def sort_messages(messages: np.ndarray):
        """By time, then pitch, note-ons (velocity > 0) before note-offs."""
        if messages.ndim == 2:
            return messages[np.lexsort((-messages[:, 3], messages[:, 2], messages[:, 0]))]
        return messages[np.lexsort((messages,))]

    @staticmethod
    def filter_messages(messages: np.ndarray, start: float = 0.):
        return messages[(messages[:, 0] if messages.ndim == 2 else messages) >= start]


class SPMuple2Messenger(SPMupleMessenger):
    """Time-based decoding of SPMuple2: every score onset gets a performed time from the previous onset's time, the local tempo and
    the notes' relative deviations; the tempo is the notes' Tempo token or re-estimated from the onsets performed so far
    (messengers.py:196-363).  The recurrence runs over the onsets of PERFORMED notes (velocity token above the zero token)."""

    def tokens_to_messages(self, tokens: np.ndarray, note_attributes: bool = True, note_on_events: bool = True,
                           note_off_events: bool = True, intermediates: Optional[SPMuple2IntermediateData] = None,
                           return_intermediates: bool = False, to_times: bool = True, sort: bool = True):
        assert to_times, "Tick messages are not supported with SPMuple2 encoding"
        tok, p = self.tokenizer, self.tokenizer.config.additional_params
        scale = 60 / self.beat_resolution
        tok._current_midi_metadata = {"tempo_scale": scale}
        from_tokens = (not p["decode_recompute_tempos"]) or p["onset_tempos"]       # tempo = the notes' own Tempo field
        re_estimate = p["decode_recompute_tempos"] and not p["onset_tempos"]        # tempo = fit over the performed onsets

        ticks = tok.compute_ticks(tokens, self.beat_resolution, compute_beat_ticks=True)["note_on"].astype(float)
        durations = tok.decode_token_type(tokens, "Duration")
        note_bpm = tok.decode_token_type(tokens, "Tempo")
        rel_dev = tok.decode_token_type(tokens, "RelOnsetDev")
        rel_held = tok.decode_token_type(tokens, "RelPerfDuration")
        performed = tokens[:, tok.vocab_types_idx["Velocity"]] != tok.zero_token

        state = intermediates if intermediates is not None else SPMuple2IntermediateData()
        tempos = state.tempos if state.tempos is not None else np.array([[state.initial_tempo, 0, 0.]])
        pairs = state.onset_pairs
        if pairs is None:                                  # an anchor one tick before the piece unless it starts late
            pairs = np.array([(0, 0, 1)]) if ticks[0] > 0 else np.array([(-1, -1 / tempos[-1, 0] * scale, 1)])

        # notes of every onset, in note order (stable), visited in increasing tick order
        order = np.argsort(ticks, kind="stable")
        starts = np.flatnonzero(np.concatenate([[True], ticks[order][1:] != ticks[order][:-1]]))
        groups = np.split(order, starts[1:])

        n = len(ticks)
        on, off = np.zeros(n), np.zeros(n)
        h = native.lib() if isinstance(tok, TokenTables) else None      # other tokenizers keep their own tempo helpers
        if h is not None:
            tempos, pairs = self._recurrence_native(h, tok, p, state, tempos, pairs, order, starts, ticks, durations, note_bpm, rel_dev,
                                                    rel_held, performed, scale, from_tokens, re_estimate, on, off)
        else:
            tempos, pairs = self._recurrence(tok, p, state, tempos, pairs, groups, ticks, durations, note_bpm, rel_dev, rel_held,
                                             performed, scale, from_tokens, re_estimate, on, off)

        messages = self._pack(tokens, on, off, note_attributes, note_on_events, note_off_events)
        if sort:
            messages = self.sort_messages(messages)
        if not return_intermediates:
            return messages
        return messages, SPMuple2IntermediateData(tempos=tempos, initial_tempo=state.initial_tempo, onset_pairs=pairs)

    @staticmethod
    def _recurrence(tok, p, state, tempos, pairs, groups, ticks, durations, note_bpm, rel_dev, rel_held, performed, scale, from_tokens,
                    re_estimate, on, off):
        """The recurrence over performed onsets, in numpy: the definition csrc/onset_times.c is tested against."""
        bpm = tempos[-1, 0]
        last_tick, last_time, last_n = pairs[-1]
        for members in groups:
            live = performed[members]
            if not live.any():
                continue
            tick, count = ticks[members[0]], len(members)
            # a chord cut by the previous chunk continues: fall back to the state before its first part.  As in the reference the
            # last rows of the caller's `tempos` / `onset_pairs` arrays are overwritten in place in that case.
            resumed = tick == tempos[-1, 1] and tick > 0
            if resumed:
                last_tick, last_time, last_n = pairs[-2]
                bpm = tempos[-2, 0]
            if from_tokens:
                bpm = (bpm * last_n + note_bpm[members].sum()) / (last_n + count) if resumed else note_bpm[members].mean()

            step = (tick - last_tick) / bpm * scale          # seconds since the previous onset at the current tempo
            played = last_time + step + rel_dev[members] * step
            if resumed:
                onset_time = (pairs[-1, 1] * last_n + played[live].sum())
                onset_time /= (last_n + count)
                pairs[-1] = np.array([tick, onset_time, last_n + count])
            else:
                onset_time = played[live].mean()
                pairs = np.concatenate([pairs, [(tick, onset_time, count)]])
            on[members] = played
            off[members] = played + rel_held[members] * (durations[members] / bpm * scale)

            if re_estimate:
                if onset_time < 2 * p["tempo_min_onset_dist"]:
                    bpm = state.initial_tempo
                else:
                    here = pairs[-1, :2]
                    bpm = tok.compute_local_tempo(distances=here - tok.filter_onsets_in_window(here, pairs[:-1, :2], index=len(pairs) - 1))
            if resumed:
                tempos[-1] = np.array([[bpm, tick, onset_time]])
                last_tick, last_time, last_n = pairs[-1]
            else:
                tempos = np.concatenate([tempos, np.array([[bpm, tick, onset_time]])])
                last_tick, last_time, last_n = tick, onset_time, count
        return tempos, pairs

    @staticmethod
    def _recurrence_native(h, tok, p, state, tempos, pairs, order, starts, ticks, durations, note_bpm, rel_dev, rel_held, performed, scale,
                           from_tokens, re_estimate, on, off):
        """The same recurrence in C (include/spb200_host.h).  The state arrays are copied into buffers with room for one row per onset;
        when the first performed onset continued the incoming state's last onset, the overwritten last rows are mirrored into the
        caller's arrays, which is what the in-place assignments of the numpy path (and of the reference) do."""
        import ctypes
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        n_groups = len(starts)
        bounds = np.ascontiguousarray(np.concatenate([starts, [len(order)]]), dtype=np.int64)
        order = np.ascontiguousarray(order, dtype=np.int64)
        nt, npair = tempos.shape[0], pairs.shape[0]
        tbuf, pbuf = np.zeros((nt + n_groups, 3)), np.zeros((npair + n_groups, 3))
        tbuf[:nt], pbuf[:npair] = tempos, pairs
        arrs = [f64(ticks), f64(durations), f64(note_bpm), f64(rel_dev), f64(rel_held), np.ascontiguousarray(performed, dtype=np.uint8)]
        table = f64(tok.tempos)
        out_t, out_p, resumed = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
        ptr = lambda a: a.ctypes.data
        rc = h.spb_host_onset_times(len(arrs[0]), *(ptr(a) for a in arrs), ptr(order), n_groups, ptr(bounds), ptr(tbuf), nt, ptr(pbuf),
                                    npair, float(scale), float(state.initial_tempo), int(bool(from_tokens)), int(bool(re_estimate)),
                                    float(p.get("tempo_min_onset_dist", 0.)), float(p.get("tempo_window", 0.)),
                                    int(p.get("tempo_min_onsets", 0)), int(bool(tok.use_tempos and p.get("use_quantized_tempos", False))),
                                    ptr(table), len(table), ptr(on), ptr(off), ctypes.byref(out_t), ctypes.byref(out_p),
                                    ctypes.byref(resumed))
        if rc == -3:
            raise IndexError("an onset continues the previous chunk but the tempo state holds a single row")
        if rc != 0:
            raise RuntimeError(f"spb_host_onset_times failed ({rc})")
        if resumed.value:
            tempos[-1], pairs[-1] = tbuf[nt - 1], pbuf[npair - 1]
        if out_t.value == nt and out_p.value == npair:     # nothing appended: the numpy path hands the caller's arrays back
            return tempos, pairs
        return tbuf[:out_t.value], pbuf[:out_p.value]

