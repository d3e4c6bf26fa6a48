#!/usr/bin/env python3
"""Developer tool (GPU box): stage times of several builds of the library in ONE process (the 1 M-sentence corpus
takes longer to synthesise than everything else), each with several run-time configurations.

    python tools/ab_multi.py [--dict synth-unidic] [--batch 1000000] [--fixed-len 0] [--check] \
        name=path/to/lib.so[:opt=val,opt=val[/opt=val,...]] ...

`--check` compares every configuration's tokens with the first configuration of the first library (bit-exact).
Appends to gpurun_out/ab_multi.txt.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from vibrato_b200 import synth  # noqa: E402
from vibrato_b200 import _native  # noqa: E402
import vibrato_b200 as vb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dict", default="synth-unidic")
    ap.add_argument("--batch", type=int, default=1000000)
    ap.add_argument("--fixed-len", type=int, default=0)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("specs", nargs="+")
    a = ap.parse_args()
    sd = synth.make_dictionary(a.dict)
    utf8, off = synth.make_corpus(sd, a.batch, seed=20260925, **({"fixed_len": a.fixed_len} if a.fixed_len else {}))
    d_utf8 = torch.from_numpy(utf8).cuda()
    d_off = torch.from_numpy(off.astype(np.int64)).cuda()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = open(os.path.join(ROOT, "gpurun_out", "ab_multi.txt"), "a")
    ref = None
    for spec in a.specs:
        name, rest = spec.split("=", 1)
        path, _, cfgs = rest.partition(":")
        _native._lib = None
        _native.SO_PATH = os.path.abspath(path)
        d = vb.SystemDictionaryBuilder.from_readers(sd.lex_csv, sd.matrix, sd.char_def, sd.unk_def)
        for cfg in (cfgs.split("/") if cfgs else [""]):
            tok = vb.Tokenizer.new(d)
            for kv in filter(None, cfg.split(",")):
                k, v = kv.split("=")
                tok.set_option(k, int(v))
            for _ in range(2):
                tok.tokenize_batch_device(d_utf8.data_ptr(), d_off.data_ptr(), a.batch, len(utf8))
            acc = None
            for _ in range(a.reps):
                res = tok.tokenize_batch_device(d_utf8.data_ptr(), d_off.data_ptr(), a.batch, len(utf8))
                ms = tok.last_stage_ms()
                acc = ms if acc is None else {k: acc[k] + v for k, v in ms.items()}
            line = f"{name:12s} {cfg:44s} " + " ".join(f"{k}={v / a.reps:6.3f}" for k, v in acc.items()) + \
                f" sum={sum(acc.values()) / a.reps:7.3f}"
            if a.check:
                d_tok_off, d_tokens, n_tok = res
                n = int(n_tok)
                toks = _read_device(int(d_tokens), n * 24)
                offs = _read_device(int(d_tok_off), (a.batch + 1) * 8)
                if ref is None:
                    ref = (toks, offs)
                    line += " check=ref"
                else:
                    line += " check=" + ("OK" if (np.array_equal(ref[0], toks) and np.array_equal(ref[1], offs)) else "MISMATCH")
            print(line, flush=True)
            out.write(line + "\n")
            out.flush()
            del tok
        del d


def _read_device(ptr, nbytes):
    """Device memory -> numpy bytes via cudaMemcpy (ctypes on the runtime torch already loaded)."""
    import ctypes as C
    rt = C.CDLL("/usr/local/cuda/lib64/libcudart.so.12")
    host = np.empty(nbytes, dtype=np.uint8)
    rc = rt.cudaMemcpy(C.c_void_p(host.ctypes.data), C.c_void_p(ptr), C.c_size_t(nbytes), C.c_int(2))
    assert rc == 0, rc
    return host


if __name__ == "__main__":
    main()
