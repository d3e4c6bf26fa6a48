#!/usr/bin/env python3
"""Developer tool (GPU box): stage times of the three connector kinds (Matrix, Raw, Dual — connector.rs:30-35) on the same
lexicon and corpus (synth-ipadic shape, 100 k sentences): the Raw / Dual cost functions of k_viterbi2 are parity-tested but
are not part of any BASELINE configuration; this records what they cost.   python tools/connector_bench.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import vibrato_b200 as vb  # noqa: E402
from vibrato_b200 import synth  # noqa: E402

sd = synth.make_dictionary("synth-ipadic")
utf8, off = synth.make_corpus(sd, 100000, seed=20260925)
right, left, cost = synth.make_bigram_files(sd, n_templates=12)
d_utf8 = torch.from_numpy(utf8).cuda()
d_off = torch.from_numpy(off.astype(np.int64)).cuda()
dicts = {
    "matrix": vb.SystemDictionaryBuilder.from_readers(sd.lex_csv, sd.matrix, sd.char_def, sd.unk_def),
    "raw (12 templates)": vb.SystemDictionaryBuilder.from_readers_with_bigram_info(sd.lex_csv, right, left, cost, sd.char_def, sd.unk_def),
    "dual (matrix + 8 raw lanes)": vb.SystemDictionaryBuilder.from_readers_with_bigram_info(sd.lex_csv, right, left, cost, sd.char_def,
                                                                                        sd.unk_def, dual_connector=True),
}
print("| connector | viterbi ms | whole step ms | sentences/s |\n|---|---:|---:|---:|")
for name, d in dicts.items():
    tok = vb.Tokenizer.new(d)
    for _ in range(2):
        tok.tokenize_batch_device(d_utf8.data_ptr(), d_off.data_ptr(), 100000, len(utf8))
    acc = None
    for _ in range(5):
        tok.tokenize_batch_device(d_utf8.data_ptr(), d_off.data_ptr(), 100000, len(utf8))
        ms = tok.last_stage_ms()
        acc = ms if acc is None else {k: acc[k] + v for k, v in ms.items()}
    tot = sum(acc.values()) / 5
    print(f"| {name} | {acc['viterbi'] / 5:.3f} | {tot:.3f} | {100000 / tot * 1e3:.3g} |", flush=True)
