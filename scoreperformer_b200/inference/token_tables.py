"""Value tables of an SPMuple vocabulary: what the rendering loop and the messengers need from the tokenizer, without `miditok`.

The reference's inference package reaches into its tokenizer for a handful of things only: field order (`vocab_types_idx`), the
special-token ids, per-field value tables (`decode_token_type`, data/tokenizers/common/octuple_m.py:371-390 and
spmuple/spmuple.py:756-775), score ticks from Bar / Position / TimeSig (`compute_ticks`, octuple_m.py:460-519), and the local-tempo
helpers of SPMuple2 (spmuple/spmuple2.py:548-593).  `TokenTables` holds exactly those tables and restates those functions in numpy,
so a rendering service can run from a checkpoint plus a small table file.  `TokenTables.from_tokenizer` lifts them out of a real
`SPMuple` / `SPMuple2` object when `miditok` is available; both the reference tokenizer and a `TokenTables` can be handed to
`messengers.SPMupleMessenger` / `generators.ScorePerformerGenerator` (duck-typed).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from types import SimpleNamespace
from typing import Dict, Optional, Tuple

import numpy as np

SPECIAL_TOKENS = ("PAD", "MASK", "SOS", "EOS")          # data/tokenizers/constants.py:4
SOS_TOKEN, EOS_TOKEN = "SOS_None", "EOS_None"
DEFAULT_TEMPO = 120                                      # miditok.constants.TEMPO
NOTE_ON_MIDI_EVENT = 144                                 # inference/messengers.py:12

_COMPOUND_BEATS = {6: 2, 9: 3, 18: 3, 12: 4, 24: 4}     # beats per bar of compound metres (octuple_m.py:508-510)


def find_closest(array: np.ndarray, values):
    """Index of the entry of the sorted `array` nearest to each value; ties go to the upper neighbour (utils/functions.py
    `find_closest`)."""
    array = np.asarray(array)
    hi = np.searchsorted(array, values, side="left")
    lo = np.maximum(hi - 1, 0)
    hi_c = np.minimum(hi, len(array) - 1)
    take_lo = (hi == len(array)) | (np.fabs(values - array[lo]) < np.fabs(values - array[hi_c]))
    return np.where(take_lo, lo, hi_c) if isinstance(hi, np.ndarray) else (int(lo) if take_lo else int(hi_c))


def is_spmuple2(tokenizer) -> bool:
    """The reference tests `isinstance(tokenizer, SPMuple2)` (generators.py:80,101,168); without `miditok` the class cannot be
    imported, so the family is read from a flag or from the class names in the MRO."""
    flag = getattr(tokenizer, "spmuple2", None)
    if flag is not None:
        return bool(flag)
    return any(c.__name__ == "SPMuple2" for c in type(tokenizer).__mro__)


@dataclass
class TokenTables:
    vocab_types_idx: Dict[str, int]                      # field name -> column of the note tuple
    sizes: Dict[str, int]                                # field name -> vocabulary size (special tokens included)
    beat_res: int                                        # samples per beat of the position grid (max of config.beat_res)
    pitch_min: int
    velocities: np.ndarray                               # value of Velocity token i (0 = unperformed note)
    duration_values: np.ndarray                          # beats of Duration token i
    tempos: np.ndarray
    time_signatures: np.ndarray                          # [n, 2]
    position_shifts: Optional[np.ndarray] = None
    rel_onset_deviations: Optional[np.ndarray] = None
    rel_performed_durations: Optional[np.ndarray] = None
    additional_params: Dict[str, object] = field(default_factory=dict)
    use_tempos: bool = True
    spmuple2: bool = True
    zero_token: int = len(SPECIAL_TOKENS)

    def __post_init__(self):
        for name in ("velocities", "duration_values", "tempos", "time_signatures", "position_shifts", "rel_onset_deviations",
                     "rel_performed_durations"):
            v = getattr(self, name)
            if v is not None:
                setattr(self, name, np.asarray(v))
        # the attribute paths the reference code walks: tokenizer.config.beat_res / .additional_params / .use_tempos
        self.config = SimpleNamespace(beat_res={(0, 4): self.beat_res}, additional_params=self.additional_params,
                                      use_tempos=self.use_tempos, pitch_range=(self.pitch_min, self.pitch_min + 88))
        self._max_beat_res = self.beat_res
        self._duration_values = self.duration_values
        self._current_midi_metadata = {}

    # ---------------------------------------------------------------- ids
    def __getitem__(self, item: Tuple[int, str]) -> int:
        """`tokenizer[dim, 'SOS_None']`, `tokenizer[dim, 'Tempo_96']` (generators.py:51-52,310)."""
        _, name = item
        kind, _, value = name.partition("_")
        if kind in SPECIAL_TOKENS:
            return SPECIAL_TOKENS.index(kind)
        table = {"Tempo": self.tempos, "Velocity": self.velocities}.get(kind)
        if table is None:
            raise KeyError(name)
        hit = np.nonzero(table == float(value))[0]
        if hit.size == 0:
            raise KeyError(name)
        return int(hit[0]) + self.zero_token

    # ---------------------------------------------------------------- values
    def decode_token_type(self, tokens: np.ndarray, token_type: str) -> np.ndarray:
        ids = tokens[:, self.vocab_types_idx[token_type]] - self.zero_token
        if token_type == "Pitch":
            return ids + self.pitch_min
        if token_type == "Velocity":
            return self.velocities[ids]
        if token_type in ("Duration", "PerfDuration"):
            return self.duration_values[ids] * self.beat_res
        if token_type == "Tempo":
            return self.tempos[ids]
        if token_type == "TimeSig":
            return self.time_signatures[ids]
        if token_type == "PositionShift":
            return self.position_shifts[ids]
        if token_type == "OnsetDev":
            return ids - 2 * self.beat_res
        if token_type == "RelOnsetDev":
            return self.rel_onset_deviations[ids]
        if token_type == "RelPerfDuration":
            return self.rel_performed_durations[ids]
        return ids

    def compute_ticks(self, tokens: np.ndarray, time_division: int = 480, compute_beat_ticks: bool = False) -> Dict[str, object]:
        """Score ticks of every note, bar and beat from the Bar / Position / TimeSig columns.  As in the reference the
        time-signature map is rebuilt from the tokens given, so it is exact for full-length or single-metre windows only."""
        ticks_per_sample = time_division / self.beat_res
        bars = self.decode_token_type(tokens, "Bar")
        positions = self.decode_token_type(tokens, "Position")

        ts_col = tokens[:, self.vocab_types_idx["TimeSig"]]
        seg_start = np.concatenate([[0], np.flatnonzero(ts_col[1:] != ts_col[:-1]) + 1])     # first note of every metre segment
        metres = self.decode_token_type(tokens[seg_start], "TimeSig")
        bar_len = time_division * 4 * metres[:, 0] / metres[:, 1]
        seg_bar = bars[seg_start]
        seg_tick = np.concatenate([[0], np.cumsum(bar_len[:-1] * np.diff(seg_bar))])

        def segment_of(n):                                 # metre segment of unit 0..n (bars or beats), clamped to the first
            return np.maximum(0, np.searchsorted(seg_bar, np.arange(n + 1), side="right") - 1)

        bar_ticks = np.concatenate([[0], np.cumsum(bar_len[segment_of(bars[-1])])])
        out = {"note_on": bar_ticks[bars] + positions * ticks_per_sample, "time_sig": (metres, seg_tick), "bar": bar_ticks}
        if compute_beat_ticks:
            beats_in_bar = metres[:, 0]                    # a view: like the reference, the returned metres carry the folded count
            folded = beats_in_bar.copy()
            for numerator, beats in _COMPOUND_BEATS.items():
                folded[beats_in_bar == numerator] = beats
            beats_in_bar[:] = folded
            beat_len = bar_len // beats_in_bar
            n_beats = int(np.sum(np.diff(np.concatenate([seg_bar, [bars[-1] + 1]])) * beats_in_bar))
            out["beat"] = np.concatenate([[0], np.cumsum(beat_len[segment_of(n_beats)])])
        return out

    def compute_position_shifts(self, score_positions: np.ndarray, onset_shift: Optional[bool] = None) -> np.ndarray:
        """Distance of every note to the previous onset (spmuple/spmuple.py:721-736)."""
        if onset_shift is None:
            onset_shift = self.additional_params["onset_position_shifts"]
        if not onset_shift:
            return np.concatenate([score_positions[:1], np.diff(score_positions)])
        onsets, inverse = np.unique(score_positions, return_inverse=True)
        order = np.sort(inverse)                            # the reference indexes by sorted onset id, not by note order
        shifts = onsets[order] - onsets[order - 1]
        neg = shifts < 0
        shifts[neg] = score_positions[neg]
        return shifts

    # ---------------------------------------------------------------- SPMuple2 local tempo
    def filter_onsets_in_window(self, onset_pair: np.ndarray, onset_pairs: np.ndarray, index: int) -> np.ndarray:
        """Earlier (tick, time) onsets that vote on the local tempo: at least `tempo_min_onset_dist` seconds back, inside
        `tempo_window` seconds, widened to the last `tempo_min_onsets` within four windows (spmuple/spmuple2.py:548-576)."""
        p = self.additional_params
        now = onset_pair[1]
        past = onset_pairs[:index]
        far = past[past[:, 1] <= now - p["tempo_min_onset_dist"]]
        if len(far) == 0:
            far = past
        win = far[far[:, 1] >= now - p["tempo_window"]]
        if len(win) < p["tempo_min_onsets"]:
            win = far[max(0, len(far) - p["tempo_min_onsets"]):]
            win = win[win[:, 1] >= now - 4 * p["tempo_window"]]
        return far if len(win) == 0 else win

    def compute_local_tempo(self, distances: np.ndarray) -> float:
        """Recency-weighted mean of tick / second ratios, floored at the slowest tempo (spmuple/spmuple2.py:578-593)."""
        dt = distances[:, 1]
        local = distances[:, 0] / dt * self._current_midi_metadata["tempo_scale"]
        w = 1 - dt / (dt.max() + 0.01)
        w /= w.sum()
        tempo = max(self.tempos[0], (w * local).sum())
        if self.use_tempos and self.additional_params["use_quantized_tempos"]:
            tempo = self.tempos[find_closest(self.tempos, tempo)]
        return tempo

    # ---------------------------------------------------------------- construction / storage
    @classmethod
    def from_tokenizer(cls, tok) -> "TokenTables":
        """Lift the tables out of a reference `SPMuple` / `SPMuple2` tokenizer (needs `miditok` in the calling process)."""
        opt = lambda name: getattr(tok, name, None)
        return cls(vocab_types_idx=dict(tok.vocab_types_idx), sizes=dict(tok.sizes), beat_res=int(max(tok.config.beat_res.values())),
                   pitch_min=int(tok.config.pitch_range[0]), velocities=np.asarray(tok.velocities),
                   duration_values=np.asarray(tok.duration_values), tempos=np.asarray(tok.tempos),
                   time_signatures=np.asarray(tok.time_signatures), position_shifts=opt("position_shifts"),
                   rel_onset_deviations=opt("rel_onset_deviations"), rel_performed_durations=opt("rel_performed_durations"),
                   additional_params=dict(tok.config.additional_params), use_tempos=bool(tok.config.use_tempos),
                   spmuple2=is_spmuple2(tok), zero_token=int(tok.zero_token))

    @classmethod
    def from_preset(cls, preset) -> "TokenTables":
        """Build the tables from a tokenizer preset of the reference (`data/tokenizers/spmuple_*.json`: a path or the parsed dict)
        without `miditok`.

        The SPMuple-specific bins restate the reference (`_create_position_shifts`, `_create_relative_onset_deviations`,
        `_create_relative_performed_durations`: spmuple/spmuple.py:653-720, spmuple/spmuple2.py:491-546; field order:
        common/octuple_m.py:295-345 + spmuple/spmuple.py:627-651) and are checked against those methods.  Velocities, durations,
        tempi and time signatures come from the third-party `miditok` (pinned 2.1.6 in the presets, not vendored in the
        reference): its published construction is restated here -- velocities `linspace(0, 127, n + 1)[1:]`, one duration per
        (beat, sample) of every `beat_res` range plus the closing one, `geomspace` / `linspace` tempi rounded to 2 decimals, time
        signatures in the order of `time_signature_range` -- and is pinned only by the vocabulary sizes it yields."""
        if isinstance(preset, str):
            import json
            with open(preset) as f:
                preset = json.load(f)
        cfg, family = preset["config"], preset.get("tokenization", "SPMuple2")
        p = dict(cfg["additional_params"])
        spm2 = family in ("SPMuple2", "SPMupleOnset", "SPMupleWindow", "SPMupleWindowRecompute")
        # what the encodings set on top of the stored parameters (spmuple/encodings.py)
        p.update({"use_position_shifts": True, "use_onset_indices": True})
        if family in ("SPMupleBeat", "SPMupleBar"):
            p.update({"rel_onset_dev": True, "rel_perf_duration": True, "bar_tempos": family == "SPMupleBar"})
        if family == "SPMupleOnset":
            p["onset_tempos"] = True
        if family == "SPMupleWindow":
            p.update({"use_quantized_tempos": True, "decode_recompute_tempos": False})
        if family == "SPMupleWindowRecompute":
            p.update({"use_quantized_tempos": p.get("use_quantized_tempos", True), "decode_recompute_tempos": True})

        ranges = {tuple(int(v) for v in k.split("_")) if isinstance(k, str) else tuple(k): int(r) for k, r in cfg["beat_res"].items()}
        res = max(ranges.values())
        durations = [(beat, pos, r) for rng, r in ranges.items() for beat in range(*rng) for pos in range(r)]
        last = max(ranges)
        durations.append((max(last), 0, ranges[last]))
        durations = [(0, 0, durations[1][-1])] + durations[1:]                   # miditok drops the zero duration, OctupleM adds one back
        duration_values = np.array([(b * r + q) / r if r > 0 else 0 for b, q, r in durations])
        velocities = np.concatenate([[0], np.linspace(0, 127, cfg["nb_velocities"] + 1, dtype=np.intc)[1:]])
        spacing = np.geomspace if cfg.get("log_tempos", False) else np.linspace
        tempos = spacing(*cfg["tempo_range"], cfg["nb_tempos"]).round(2)
        metres = np.array([(int(n), int(d)) for d, beats in cfg["time_signature_range"].items() for n in beats])
        n_positions = int(max(np.ceil(4 * metres[:, 0] / metres[:, 1]))) * res

        shifts = np.concatenate([np.arange(0, 2 * res, 1), np.arange(2 * res, 4 * res, 2), np.arange(4 * res, 8 * res, 8),
                                 np.arange(8 * res, 16 * res + 1, 16)])
        rel_dev, rel_dur = _relative_bins(p["nb_onset_devs"], p["nb_perf_durations"], spm2)

        names = ["Bar", "Position", "Pitch", "Velocity", "Duration"] + (["Tempo"] if cfg["use_tempos"] else []) \
            + (["TimeSig"] if cfg["use_time_signatures"] else []) + (["Program"] if cfg.get("use_programs") else []) \
            + ["PositionShift", "NotesInOnset", "PositionInOnset"] \
            + ["RelOnsetDev" if p["rel_onset_dev"] else "OnsetDev", "RelPerfDuration" if p["rel_perf_duration"] else "PerfDuration"]
        counts = {"Bar": p["max_bar_embedding"], "Position": n_positions, "Pitch": cfg["pitch_range"][1] - cfg["pitch_range"][0],
                  "Velocity": len(velocities), "Duration": len(durations), "Tempo": len(tempos), "TimeSig": len(metres),
                  "Program": len(cfg.get("programs", [])), "PositionShift": len(shifts), "NotesInOnset": p["max_notes_in_onset"],
                  "PositionInOnset": p["max_notes_in_onset"], "RelOnsetDev": len(rel_dev), "OnsetDev": 4 * res + 1,
                  "RelPerfDuration": len(rel_dur), "PerfDuration": len(durations)}
        zero = len(cfg.get("special_tokens", SPECIAL_TOKENS))
        return cls(vocab_types_idx={k: i for i, k in enumerate(names)}, sizes={k: counts[k] + zero for k in names}, beat_res=res,
                   pitch_min=int(cfg["pitch_range"][0]), velocities=velocities, duration_values=duration_values, tempos=tempos,
                   time_signatures=metres, position_shifts=shifts, rel_onset_deviations=rel_dev, rel_performed_durations=rel_dur,
                   additional_params=p, use_tempos=bool(cfg["use_tempos"]), spmuple2=spm2, zero_token=zero)

    _ARRAYS = ("velocities", "duration_values", "tempos", "time_signatures", "position_shifts", "rel_onset_deviations",
               "rel_performed_durations")

    def save(self, path: str) -> None:
        import json
        arrays = {k: getattr(self, k) for k in self._ARRAYS if getattr(self, k) is not None}
        meta = dict(vocab_types_idx=self.vocab_types_idx, sizes=self.sizes, beat_res=self.beat_res, pitch_min=self.pitch_min,
                    additional_params=self.additional_params, use_tempos=self.use_tempos, spmuple2=self.spmuple2,
                    zero_token=self.zero_token)
        np.savez(path, meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **arrays)

    @classmethod
    def load(cls, path: str) -> "TokenTables":
        import json
        with np.load(path) as z:
            meta = json.loads(bytes(z["meta"]).decode())
            return cls(**meta, **{k: z[k] for k in cls._ARRAYS if k in z.files})


def _relative_bins(n_dev: int, n_dur: int, spm2: bool):
    """Bins of the RelOnsetDev / RelPerfDuration fields: piecewise linear near zero / one, geometric in the tails, rounded to four
    decimals; onset deviations mirrored to negative values (spmuple/spmuple.py:668-720 for SPMuple, spmuple/spmuple2.py:491-546 for
    SPMuple2 -- the two families place their breakpoints differently)."""
    lin, steps = np.linspace, np.arange
    if spm2:
        q = (n_dev - 1) // 10
        log32, log43 = np.log(3 / 2) / np.log(2), np.log(4 / 3) / np.log(2)
        dev = [lin(0, 1 / 20, q + 1), lin(1 / 20, 1 / 10, q + 1)[1:], lin(1 / 10, 1 / 6, q + 1)[1:],
               (2 ** (steps(q + 1) / q) * 1 / 6)[1:], (2 ** (log32 * steps(q // 2 + 1) / q * 2) * 1 / 3)[1:],
               (2 ** (log32 * steps(q // 4 + 1) / q * 4) * 1 / 2)[1:], (2 ** (log43 * steps(q // 8 + 1) / q * 8) * 3 / 4)[1:],
               (2 ** (steps(q // 8 + 1) / q * 8))[1:]]
        d = (n_dur - 1) // 5
        dur = [lin(1 / 10, 1 / 3, d + 1), lin(1 / 3, 4 / 5, 2 * d + 1)[1:], lin(4 / 5, 1., d + 1)[1:], lin(1.0, 5 / 4, d // 2 + 1)[1:],
               lin(5 / 4, 3 / 2, d // 4 + 1)[1:], (2 ** (4 * steps(d // 4 + 1) / d) * 3 / 2)[1:]]
    else:
        q = (n_dev - 1) // 8
        dev = [lin(0.0, 1 / 24, q + 1), lin(1 / 24, 1 / 8, q + 1)[1:], lin(1 / 8, 1 / 3, q + 1)[1:], lin(1 / 3, 3 / 5, q // 2 + 1)[1:],
               lin(3 / 5, 1.0, q // 4 + 1)[1:], (2 ** (8 * steps(q // 4 + 1) / q))[1:]]
        d = (n_dur - 1) // 4
        dur = [lin(1 / 10, 2 / 5, d + 1), lin(2 / 5, 2 / 3, d + 1)[1:], lin(2 / 3, 1.0, d + 1)[1:], lin(1.0, 5 / 4, d // 2 + 1)[1:],
               lin(5 / 4, 3 / 2, d // 4 + 1)[1:], (2 ** (4 * steps(d // 4 + 1) / d) * 3 / 2)[1:]]
    dev = np.round(np.concatenate(dev), 4)
    return np.sort(np.concatenate([-dev[1:], dev])), np.round(np.concatenate(dur), 4)
