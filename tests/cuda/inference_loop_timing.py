"""Real-time rendering loop on the GPU: the reference's generator next to scoreperformer_b200.inference (SURVEY §8 f3 measurement).

    python tests/cuda/inference_loop_timing.py [--notes 300] [--window 0.5] [--fake]

Three arms render the same synthetic piece (oracle/inference_cases.make_piece: chords of 1-4 notes) greedily in consecutive time
windows:
  A  the UNMODIFIED reference generator + messenger (baseline/_ref) around the UNMODIFIED reference model on the host CPU (fp32, all
     cores; the first `--ref-notes` notes of the piece).  On the GPU the reference's cached decoding aborted with an illegal memory
     access inside its eager forward on this image (first attempt of this script, profiles/r02_reference_gpu_decode_abort.txt), so the CPU is
     where the reference arm runs;
  B  the reference generator + messenger around this repo's CUDA decoder (`unmask_tokens` with the reference's cache contract);
  C  scoreperformer_b200.inference.ScorePerformerGenerator + SPMuple2Messenger around the same CUDA decoder;
  D  (--lookahead N) C with several chords per decoder call; must render what C renders.
B and C hand the decoder the same windows, so their tokens and messages must be IDENTICAL (asserted); A is the timing of the
reference end to end (its encoders see the shorter piece, so its tokens are not compared here -- tests/test_decode_gpu.py does that).  Wall clock per rendered note with a synchronize on both sides, after a warm-up rendering.
`--fake` replaces the model by oracle/inference_cases.FakeDecoder (CPU; for checking the script itself).
Test / measurement infrastructure: uses oracle/ and baseline/_ref, never imported by the package.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
for cand in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
    if os.path.isdir(os.path.join(cand, "scoreperformer")):
        os.environ["SPB200_REFERENCE_ROOT"] = cand
        break

import gen_inference_golden as ref  # noqa: E402  (installs the import stubs, imports the reference inference classes)
import inference_cases as cases  # noqa: E402
from scoreperformer_b200.inference import ScorePerformerGenerator, SPMuple2IntermediateData, SPMuple2Messenger, TokenTables  # noqa: E402

RENDERED = [3, 5, 10, 11]


def segments(piece):
    tick = (piece[:, 0] - 4) * 32 + (piece[:, 1] - 4)
    _, onset = np.unique(tick, return_inverse=True)
    return piece[:, 0].copy(), 4 + tick // 8, 4 + onset


def encode(model, piece, device):
    bars, beats, onsets = (torch.from_numpy(np.ascontiguousarray(s))[None].to(device) for s in segments(piece))
    perf = torch.from_numpy(piece)[None].to(device)
    mask = torch.ones(perf.shape[:2], dtype=torch.bool, device=device)
    with torch.inference_mode():
        enc = model.forward_encoders(perf=perf, perf_mask=mask, score=perf[..., :10].contiguous(), score_mask=mask, bars=bars, beats=beats,
                                     onsets=onsets, deadpan_mask=torch.zeros(1, dtype=torch.bool, device=device), compute_loss=False)
    pad = lambda e: torch.cat([e[0, :1], e[0], e[0, -1:]]).float().clone()
    return pad(enc.score_embeddings), pad(enc.perf_embeddings)


def render(generator_cls, messenger, tokenizer, model, piece, emb, interm_cls, top_k, window, device, **extra):
    gen = generator_cls(model, cases.make_dataset(tokenizer, [piece]), cases.make_collator(), messenger, device=device)
    notes = np.concatenate([np.full_like(piece[:1], 2), piece, np.full_like(piece[:1], 3)])
    notes[1:-1, RENDERED] = 1
    pd = gen.perf_data
    pd.perf_seq, pd.notes = piece, torch.from_numpy(notes).to(device)
    pd.context, pd.embeddings = (None, None) if emb is None else (emb[0].clone(), emb[1].clone())
    if emb is None:                                    # --fake: any embeddings do
        pd.context, pd.embeddings = cases.note_embeddings(len(notes)), cases.note_embeddings(len(notes)) + 100.
    pd.intermediates = interm_cls(initial_tempo=96.)
    if device != "cpu":
        torch.cuda.synchronize()
    t0, t, windows, messages = time.perf_counter(), 0., 0, []
    while not pd.reached_eos and windows < 5000:
        _, m = gen.generate_performance_notes(start_time=t, time_window=window, filter_logits_fn=top_k, filter_kwargs={"k": 1}, **extra)
        if len(m):
            messages.append(np.asarray(m))
        t += window
        windows += 1
    if device != "cpu":
        torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    assert pd.reached_eos
    calls = len(model.perf_decoder.log) if hasattr(model.perf_decoder, "log") else None
    return dict(tokens=pd.gen_seq.cpu().numpy(), messages=np.concatenate(messages), wall=wall, windows=windows, calls=calls)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--notes", type=int, default=300)
    ap.add_argument("--window", type=float, default=0.5)
    ap.add_argument("--ref-notes", type=int, default=60)
    ap.add_argument("--lookahead", type=int, default=0, help="arm D: this package's loop with lookahead_notes=N (0 = skip)")
    ap.add_argument("--fake", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    device = "cpu" if args.fake else "cuda"
    tables = TokenTables(**cases.table_kwargs())
    ref_tok = ref._ref_tokenizer(ref.SPMuple2)
    warm, piece = cases.make_piece(30, 41), cases.make_piece(args.notes, 42)

    arms = {}
    if args.fake:
        from tests.test_inference_host import Attn, Caches, Inter
        from scoreperformer_b200.modules.sampling import top_k
        model_b = cases.make_model(cases.FakeDecoder(ref.REF_CACHES))
        model_c = cases.make_model(cases.FakeDecoder((Caches, Inter, Attn)))
        plan = [("B", ref.ScorePerformerGenerator, ref.SPMuple2Messenger(ref_tok), ref_tok, model_b, ref.SPMuple2IntermediateData, top_k),
                ("C", ScorePerformerGenerator, SPMuple2Messenger(tables), tables, model_c, SPMuple2IntermediateData, top_k)]
        embs = {"B": (None, None), "C": (None, None)}
    else:
        from tests import parity
        from scoreperformer_b200.modules.sampling import top_k
        from scoreperformer.modules.sampling import top_k as ref_top_k
        import ref_shim
        model = parity.build_model(dropout=False, device="cuda").eval()
        embs = {"B": (encode(model, warm, device), encode(model, piece, device))}
        embs["C"] = embs["B"]
        plan = [("B", ref.ScorePerformerGenerator, ref.SPMuple2Messenger(ref_tok), ref_tok, model, ref.SPMuple2IntermediateData, top_k),
                ("C", ScorePerformerGenerator, SPMuple2Messenger(tables), tables, model, SPMuple2IntermediateData, top_k)]

    for name, gcls, msgr, tok, mdl, icls, tk in plan:
        render(gcls, msgr, tok, mdl, warm, embs[name][0], icls, tk, args.window, device)                 # warm-up
        if hasattr(mdl.perf_decoder, "log"):
            mdl.perf_decoder.log.clear()
        arms[name] = render(gcls, msgr, tok, mdl, piece, embs[name][1], icls, tk, args.window, device)

    if args.lookahead > 0:                             # arm D: several chords per decoder call
        name, gcls, msgr, tok, mdl, icls, tk = plan[-1]
        render(gcls, msgr, tok, mdl, warm, embs[name][0], icls, tk, args.window, device, lookahead_notes=args.lookahead)
        arms["D"] = render(gcls, msgr, tok, mdl, piece, embs[name][1], icls, tk, args.window, device, lookahead_notes=args.lookahead)
        assert np.array_equal(arms["D"]["tokens"], arms["C"]["tokens"]), "lookahead changed the rendering"

    if not args.fake and args.ref_notes > 0:          # arm A on the host cores
        torch.set_num_threads(os.cpu_count() or 1)
        ref_model, _ = ref_shim.build_reference_model(seed=23)
        ref_model.load_state_dict({k: v.cpu() for k, v in model.state_dict().items()}, strict=True)
        ref_model.eval()
        short = piece[:args.ref_notes]
        arms["A"] = render(ref.ScorePerformerGenerator, ref.SPMuple2Messenger(ref_tok), ref_tok, ref_model, short,
                           encode(ref_model, short, "cpu"), ref.SPMuple2IntermediateData, ref_top_k, args.window, "cpu")
        arms["A"]["notes"] = args.ref_notes

    b, c = arms["B"], arms["C"]
    assert np.array_equal(b["tokens"], c["tokens"]), "reference loop and this loop rendered different tokens on the same decoder"
    assert np.array_equal(b["messages"], c["messages"]), "messages differ"
    n = args.notes
    lines = [f"rendering loop on {torch.cuda.get_device_name(0) if device != 'cpu' else 'cpu (fake decoder)'}: {n} notes, "
             f"{args.window} s windows, greedy; B and C: tokens and {len(c['messages'])} messages identical"]
    label = {"A": f"reference generator + reference model, CPU fp32, {os.cpu_count()} threads", "B": "reference generator + CUDA decoder of this repo",
             "C": "scoreperformer_b200.inference + CUDA decoder of this repo",
             "D": f"the same with lookahead_notes={args.lookahead}"}
    for name in sorted(arms):
        r = arms[name]
        k = r.get("notes", n)
        lines.append(f"  {name}  {label[name]:62s} {k:4d} notes {r['wall'] * 1e3:9.1f} ms  {r['wall'] / k * 1e3:7.3f} ms/note  "
                     f"{k / r['wall']:8.1f} notes/s  ({r['windows']} windows)")
    text = "\n".join(lines)
    print(text)
    if args.out:
        with open(args.out, "w") as f:
            f.write(text + "\n" + json.dumps({k: {"wall_s": v["wall"], "windows": v["windows"]} for k, v in arms.items()}) + "\n")


if __name__ == "__main__":
    main()
