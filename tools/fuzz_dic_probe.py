#!/usr/bin/env python3
"""Robustness probe (run under compute-sanitizer): byte-flipped `.dic` streams that the loader still accepts must
not make any kernel read or write out of bounds.   python tools/fuzz_dic_probe.py [mutations per variant]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vibrato_b200 as vb  # noqa: E402
from vibrato_b200 import synth  # noqa: E402

n_mut = int(sys.argv[1]) if len(sys.argv) > 1 else 100
sd = synth.make_dictionary("synth-tiny")
right, left, cost = synth.make_bigram_files(sd, n_templates=12)
build = vb.SystemDictionaryBuilder
variants = {
    "matrix": build.from_readers(sd.lex_csv, sd.matrix, sd.char_def, sd.unk_def),
    "raw": build.from_readers_with_bigram_info(sd.lex_csv, right, left, cost, sd.char_def, sd.unk_def),
    "dual": build.from_readers_with_bigram_info(sd.lex_csv, right, left, cost, sd.char_def, sd.unk_def, dual_connector=True),
}
variants["matrix"].reset_user_lexicon_from_reader(synth.make_user_csv(sd, 50))
utf8, off = synth.make_corpus(sd, 64, seed=3, log_uniform=(1, 80), unk_frac=0.2, space_frac=0.05)
rng = np.random.default_rng(20260923)
for name, d in variants.items():
    good = bytes(d.write())
    ran = refused = rejected_later = 0
    for k in range(n_mut):
        b = bytearray(good)
        for _ in range(int(rng.integers(1, 4))):
            b[int(rng.integers(21, len(b)))] = int(rng.integers(0, 256))
        try:
            dd = vb.Dictionary.read(bytes(b))
        except vb.VibratoError:
            refused += 1
            continue
        try:
            tok = vb.Tokenizer.new(dd).ignore_space(True).max_grouping_len(8)
            tok.output_mode("detail")
            res = tok.tokenize_batch(utf8=utf8, byte_offsets=off)
            res.text()
            ran += 1
        except vb.VibratoError:
            rejected_later += 1
    print(f"{name}: {ran} mutated dictionaries tokenised, {refused} refused by the reader, {rejected_later} refused later", flush=True)
