#!/usr/bin/env python3
"""Developer tool (GPU box, `gpurun --gpus N`): the multi-device tokenizer INSIDE the library (vbt_tokenizer_new_multi)
on 1, 2, 4, ... devices of one process — host batches (one merged result) and device-resident batches (token records
gathered on devices[0] with ncclSend / ncclRecv) — plus the latency of very small batches on one device, which is
what an unmodified `Worker::reset_sentence` / `tokenize` loop pays per sentence.

    python tools/multi_bench.py [--dict synth-unidic] [--batch 1000000] [--out gpurun_out/multi_bench.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import vibrato_b200 as vb  # noqa: E402
from vibrato_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dict", default="synth-unidic")
    ap.add_argument("--batch", type=int, default=1000000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "multi_bench.json"))
    a = ap.parse_args()
    n_gpu = torch.cuda.device_count()
    sd = synth.make_dictionary(a.dict)
    utf8, off = synth.make_corpus(sd, a.batch, seed=20260925)
    d = vb.SystemDictionaryBuilder.from_readers(sd.lex_csv, sd.matrix, sd.char_def, sd.unk_def)
    h_utf8 = torch.from_numpy(utf8).pin_memory().numpy()
    h_off = torch.from_numpy(off.astype(np.int64)).pin_memory().numpy().view(np.uint64)
    torch.cuda.set_device(0)
    d_utf8 = torch.from_numpy(utf8).cuda()
    d_off = torch.from_numpy(off.astype(np.int64)).cuda()
    out = {"dict": a.dict, "batch": a.batch, "gpus_visible": n_gpu, "multi": [], "small_batches": []}
    ref_tokens = None
    for n in [x for x in (1, 2, 4, 8) if x <= n_gpu]:
        t0 = time.perf_counter()
        tok = vb.Tokenizer.new(d, devices=list(range(n)))
        info = tok.describe()
        setup_s = time.perf_counter() - t0
        for _ in range(2):
            r = tok.tokenize_batch(utf8=h_utf8, byte_offsets=h_off)
            nt = r.n_tokens
            if ref_tokens is None:
                ref_tokens = (r.tok_offsets.copy(), r.tokens.copy())
            else:  # bit-identical to the single-device result
                assert np.array_equal(r.tok_offsets, ref_tokens[0]) and r.tokens.tobytes() == ref_tokens[1].tobytes()
            del r  # frees the result without copying it out (close() would)
        torch.cuda.synchronize()
        host_s = 0.0
        for _ in range(a.reps):
            t = time.perf_counter()
            r = tok.tokenize_batch(utf8=h_utf8, byte_offsets=h_off)
            host_s += time.perf_counter() - t
            del r
        host_ms = host_s / a.reps * 1e3
        for _ in range(2):
            tok.tokenize_batch_device(d_utf8.data_ptr(), d_off.data_ptr(), a.batch, len(utf8))
        t = time.perf_counter()
        for _ in range(a.reps):
            _, _, ntd = tok.tokenize_batch_device(d_utf8.data_ptr(), d_off.data_ptr(), a.batch, len(utf8))
        dev_ms = (time.perf_counter() - t) / a.reps * 1e3
        assert ntd == nt
        row = {"devices": n, "describe": info, "setup_s": round(setup_s, 2), "host_e2e_ms": round(host_ms, 3),
               "host_e2e_sent_per_s": round(a.batch / host_ms * 1e3), "device_gather_ms": round(dev_ms, 3),
               "device_gather_sent_per_s": round(a.batch / dev_ms * 1e3), "tokens": int(nt),
               "timing": "host wall clock around the blocking C-ABI call (it synchronises every device), mean of %d" % a.reps}
        print(json.dumps(row), flush=True)
        out["multi"].append(row)
        del tok
    # small batches on one device: what a per-sentence Worker loop costs
    tok = vb.Tokenizer.new(d)
    sents = []
    for i in range(256):
        sents.append(bytes(utf8[int(off[i]):int(off[i + 1])]).decode("utf-8"))
    for n in (1, 8, 32, 256):
        u8, o = vb.Tokenizer.pack(sents[:n])
        for _ in range(20):
            tok.tokenize_batch(utf8=u8, byte_offsets=o)
        t = time.perf_counter()
        reps = 200
        for _ in range(reps):
            tok.tokenize_batch(utf8=u8, byte_offsets=o)
        us = (time.perf_counter() - t) / reps * 1e6
        row = {"n_sent": n, "latency_us": round(us, 1), "sent_per_s": round(n / us * 1e6)}
        print(json.dumps(row), flush=True)
        out["small_batches"].append(row)
    w = tok.new_worker()
    t = time.perf_counter()
    for i in range(200):
        w.reset_sentence(sents[i % 256])
        w.tokenize()
    us = (time.perf_counter() - t) / 200 * 1e6
    out["worker_loop"] = {"latency_us_per_sentence": round(us, 1), "sent_per_s": round(1e6 / us),
                          "what": "Worker::reset_sentence + tokenize per sentence through the Python mirror (tokenize/src/main.rs:78-81)"}
    print(json.dumps(out["worker_loop"]), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
