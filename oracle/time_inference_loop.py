"""TEST INFRASTRUCTURE ONLY -- host bookkeeping of the rendering loop, reference vs scoreperformer_b200.inference, on CPU tensors.

    python oracle/time_inference_loop.py           (authoring container: needs /root/reference)

Both generators render the same synthetic piece (oracle/inference_cases.py, scenario `chords_ctx48`) with the same stand-in decoder;
the time spent inside the decoder and inside the messenger is subtracted, what is left is the loop's own bookkeeping per decoder
call.  On a GPU every device-tensor expression of the reference's loop additionally costs a launch and a blocking read; this
script only shows the part visible without one.
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import gen_inference_golden as ref  # noqa: E402  (installs the stubs, imports the reference classes)
import inference_cases as cases  # noqa: E402
from scoreperformer_b200.inference import ScorePerformerGenerator, SPMuple2IntermediateData, SPMuple2Messenger, TokenTables  # noqa: E402
from tests.test_inference_host import Attn, Caches, Inter  # noqa: E402


class Clock:
    def __init__(self):
        self.t = 0.

    def wrap(self, fn):
        def timed(*a, **k):
            t0 = time.perf_counter()
            try:
                return fn(*a, **k)
            finally:
                self.t += time.perf_counter() - t0
        return timed


def measure(generator_cls, messenger_cls, tok, cache_classes, inter_cls, reps=5):
    best = None
    for _ in range(reps):
        inner, outer = Clock(), Clock()
        orig_loop = generator_cls.generate_performance_notes
        generator_cls.generate_performance_notes = outer.wrap(orig_loop)
        orig_unmask, orig_msg = cases.FakeDecoder.unmask_tokens, messenger_cls.tokens_to_messages
        cases.FakeDecoder.unmask_tokens = inner.wrap(orig_unmask)
        messenger_cls.tokens_to_messages = inner.wrap(orig_msg)
        try:
            windows, _ = cases.run_scenario("chords_ctx48", generator_cls, messenger_cls, tok, cache_classes, inter_cls)
            total = outer.t
        finally:
            generator_cls.generate_performance_notes = orig_loop
            cases.FakeDecoder.unmask_tokens, messenger_cls.tokens_to_messages = orig_unmask, orig_msg
        calls = sum(len(w["calls"]) for w in windows)
        own = (total - inner.t) / calls * 1e6
        best = own if best is None else min(best, own)
    return best, calls


if __name__ == "__main__":
    r, n = measure(ref.ScorePerformerGenerator, ref.SPMuple2Messenger, ref._ref_tokenizer(ref.SPMuple2), ref.REF_CACHES,
                   ref.SPMuple2IntermediateData)
    m, _ = measure(ScorePerformerGenerator, SPMuple2Messenger, TokenTables(**cases.table_kwargs()), (Caches, Inter, Attn),
                   SPMuple2IntermediateData)
    print(f"loop bookkeeping per decoder call ({n} calls, CPU tensors, best of 5): reference {r:.0f} us, scoreperformer_b200 {m:.0f} us")
