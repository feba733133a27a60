"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/inference_*.npz by running the UNMODIFIED reference inference package.

    python oracle/gen_inference_golden.py          (authoring container, where /root/reference exists)

What runs here is the reference's own `ScorePerformerGenerator`, `SPMupleMessenger`, `SPMuple2Messenger` and the reference
tokenizer METHODS they call (`OctupleM.compute_ticks`, `SPMuple.decode_token_type`, `SPMuple.compute_position_shifts`,
`SPMuple2.filter_onsets_in_window`, `SPMuple2.compute_local_tempo`, `utils.find_closest`).  `miditok` is not installed, so the
tokenizer OBJECT cannot be constructed the normal way: `RefTok` subclasses the reference class, skips its `__init__`, and sets
the value tables of oracle/inference_cases.py as attributes -- every method that computes anything is the reference's.  The
decoder is oracle/inference_cases.FakeDecoder (the loop, not the network, is what these goldens pin).
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_shim  # noqa: E402

ref_shim.install_stubs()

import inference_cases as cases  # noqa: E402
from scoreperformer.data.tokenizers import SPMuple, SPMuple2  # noqa: E402
from scoreperformer.inference.generators import ScorePerformerGenerator  # noqa: E402
from scoreperformer.inference.messengers import SPMupleMessenger, SPMuple2Messenger, SPMuple2IntermediateData  # noqa: E402
from scoreperformer.models.scoreperformer import TupleTransformerCaches  # noqa: E402
from scoreperformer.modules.transformer import AttentionIntermediates, TransformerIntermediates  # noqa: E402
from scoreperformer.utils import find_closest  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
REF_CACHES = (TupleTransformerCaches, TransformerIntermediates, AttentionIntermediates)


def _ref_tokenizer(base, **params):
    kw = cases.table_kwargs(**params)

    class RefTok(base):
        def __init__(self):                               # the tables, nothing else; no miditok vocabulary is built
            self.vocab_types_idx = kw["vocab_types_idx"]
            self.special_tokens = ["PAD", "MASK", "SOS", "EOS"]
            self.config = SimpleNamespace(beat_res={(0, 4): kw["beat_res"]}, additional_params=kw["additional_params"],
                                          use_tempos=True, pitch_range=(kw["pitch_min"], kw["pitch_min"] + 88))
            self._max_beat_res = kw["beat_res"]
            self.velocities, self.tempos = kw["velocities"], kw["tempos"]
            self._duration_values = kw["duration_values"]
            self.time_signatures = [tuple(x) for x in kw["time_signatures"].tolist()]
            self.position_shifts = kw["position_shifts"]
            self.rel_onset_deviations = kw["rel_onset_deviations"]
            self.rel_performed_durations = kw["rel_performed_durations"]
            self._current_midi_metadata = {}

        sizes = property(lambda self: kw["sizes"])

        def __getitem__(self, item):                      # vocabulary lookup (miditok's job): specials and Tempo_<v>
            kind, _, value = item[1].partition("_")
            if kind in self.special_tokens:
                return self.special_tokens.index(kind)
            return int(np.nonzero(self.tempos == float(value))[0][0]) + self.zero_token

    return RefTok()


def generator_goldens():
    out = {}
    for name, (_, params, _, _, _, kind) in cases.SCENARIOS.items():
        tok = _ref_tokenizer(SPMuple2, **params)
        windows, final = cases.run_scenario(name, ScorePerformerGenerator, SPMuple2Messenger, tok, REF_CACHES, SPMuple2IntermediateData)
        out[f"{name}/n_windows"] = np.int64(len(windows))
        for i, w in enumerate(windows):
            for k, v in w.items():
                out[f"{name}/w{i}/{k}"] = v
        for k, v in final.items():
            out[f"{name}/final/{k}"] = v
        print(name, len(windows), "windows,", int(sum(len(w["calls"]) for w in windows)), "decoder calls,",
              final["gen_seq"].shape[0], "notes kept")
    np.savez_compressed(os.path.join(GOLDEN_DIR, "inference_generator.npz"), **out)


MESSENGER_CASES = {
    # name: (tokenizer family, messenger, table params, piece seed, metre change)
    "spm2_refit": ("spm2", {}, 11, False),
    "spm2_refit_raw": ("spm2", dict(use_quantized_tempos=False, tempo_min_onsets=3, tempo_window=2.), 12, False),
    "spm2_token_tempo": ("spm2", dict(decode_recompute_tempos=False), 13, False),
    "spm2_onset_tempo": ("spm2", dict(onset_tempos=True), 14, True),
    "spm_beat": ("spm", {}, 15, False),
    "spm_bar_abs": ("spm", dict(bar_tempos=True, use_position_shifts=False, onset_position_shifts=True), 16, True),
    "spm_plain_shift": ("spm", dict(use_position_shifts=False, onset_position_shifts=False), 17, False),
}


def messenger_goldens():
    out = {}
    for name, (family, params, seed, metre) in MESSENGER_CASES.items():
        tok = _ref_tokenizer(SPMuple2 if family == "spm2" else SPMuple, **params)
        msgr = (SPMuple2Messenger if family == "spm2" else SPMupleMessenger)(tok)
        piece, cuts = cases.random_chunks(seed, metre_change=metre)
        state, lo = None, 0
        for i, hi in enumerate(cuts):
            chunk = piece[lo:hi]
            if family == "spm" and i % 3 == 2:           # tick messages, unsorted, without passing through times
                out[f"{name}/c{i}/ticks"] = msgr.tokens_to_messages(chunk.copy(), intermediates=state, to_times=False, sort=False)
            messages, state = msgr.tokens_to_messages(chunk.copy(), intermediates=state, return_intermediates=True)
            out[f"{name}/c{i}/messages"] = messages
            out[f"{name}/c{i}/onsets"] = msgr.tokens_to_messages(chunk.copy(), note_attributes=False, note_off_events=False,
                                                                 intermediates=None, sort=False)
            lo = hi
        out[f"{name}/tempos"] = np.array(state.tempos, dtype=np.float64)
        if family == "spm2":
            out[f"{name}/pairs"] = np.array(state.onset_pairs, dtype=np.float64)
        whole = msgr.tokens_to_messages(piece.copy())
        out[f"{name}/whole"] = whole
        print(name, len(cuts), "chunks,", whole.shape[0], "messages")

    # the tokenizer functions on their own
    tok = _ref_tokenizer(SPMuple2)
    for metre in (False, True):
        piece = cases.make_piece(200, 31, metre)
        ticks = tok.compute_ticks(piece, 8, compute_beat_ticks=True)
        tag = f"ticks{int(metre)}"
        out[f"{tag}/note_on"], out[f"{tag}/bar"], out[f"{tag}/beat"] = ticks["note_on"], ticks["bar"], ticks["beat"]
        out[f"{tag}/metres"], out[f"{tag}/metre_ticks"] = ticks["time_sig"]
        out[f"{tag}/shifts_onset"] = tok.compute_position_shifts(ticks["note_on"].copy(), onset_shift=True)
        out[f"{tag}/shifts_plain"] = tok.compute_position_shifts(ticks["note_on"].copy(), onset_shift=False)
    rng = np.random.default_rng(5)
    probe = np.concatenate([rng.uniform(20, 290, 200), tok.tempos[:5], (tok.tempos[:5] + tok.tempos[1:6]) / 2])
    out["closest/probe"], out["closest/index"] = probe, find_closest(tok.tempos, probe.copy())
    out["closest/scalars"] = np.array([find_closest(tok.tempos, float(v)) for v in probe[:40]])
    np.savez_compressed(os.path.join(GOLDEN_DIR, "inference_messenger.npz"), **out)


def encode_goldens():
    out = {}
    for name in cases.ENCODE_CASES:
        r = cases.run_encode_case(name, ScorePerformerGenerator, SPMuple2Messenger, _ref_tokenizer(SPMuple2))
        for k, v in r.items():
            out[f"{name}/{k}"] = v
        print(name, len(r["windows"]), "windows served,", len(r["calls"]), "encoder calls,", r["score"].shape[0], "embedding rows")
    np.savez_compressed(os.path.join(GOLDEN_DIR, "inference_encode.npz"), **out)


def vocabulary_goldens():
    """The SPMuple-specific bins from the reference's own constructors (the miditok-made tables cannot be produced here)."""
    out = {}
    for family, base in (("spm", SPMuple), ("spm2", SPMuple2)):
        for n_dev, n_dur, res in ((161, 81, 16), (81, 41, 8)):
            stub = base.__new__(base)
            stub._max_beat_res = res
            stub.config = SimpleNamespace(additional_params={"nb_onset_devs": n_dev, "nb_perf_durations": n_dur})
            tag = f"{family}/{n_dev}_{n_dur}_{res}"
            out[f"{tag}/position_shifts"] = stub._create_position_shifts()
            out[f"{tag}/rel_onset_deviations"] = stub._create_relative_onset_deviations()
            out[f"{tag}/rel_performed_durations"] = stub._create_relative_performed_durations()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "inference_vocab.npz"), **out)
    print("vocabulary bins:", {k: v.shape[0] for k, v in out.items() if "161" in k})


if __name__ == "__main__":
    vocabulary_goldens()
    encode_goldens()
    generator_goldens()
    messenger_goldens()
