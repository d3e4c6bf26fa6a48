#!/usr/bin/env python3
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck), checked against the oracle;
lives under tests/ because only tests may use oracle/.   compute-sanitizer python tests/probes/sanitize_probe.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import vibrato_b200 as vb  # noqa: E402
from vibrato_b200 import synth  # noqa: E402
from oracle import vibrato_oracle as vo  # noqa: E402

g = json.load(open(os.path.join(ROOT, "tests", "golden", "vibrato_fixture.json"), encoding="utf-8"))
r = g["resources"]
d = vb.SystemDictionaryBuilder.from_readers(r["lex.csv"], r["matrix.def"], r["char.def"], r["unk.def"])
d.reset_user_lexicon_from_reader(r["user.csv"])
sents = [c["input"] for c in g["tokenizer_cases"]] * 20 + ["X" * 300, "", "0123456789" * 40]
for ign in (0, 1):
    tok = vb.Tokenizer.new(d).ignore_space(bool(ign)).max_grouping_len(24 if ign else 0)
    for lanes, smem, chunk in ((16, 0, 0), (8, 0, 100), (32, 0, 64)):
        tok.set_option("lanes_per_sentence", lanes)
        tok.set_option("chunk_sentences", chunk)
        res = tok.tokenize_batch(sents)
        print("ignore_space", ign, "lanes", lanes, "smem", smem, "chunk", chunk, "tokens", res.n_tokens, flush=True)
sd = synth.make_dictionary("synth-tiny")
d2 = vb.SystemDictionaryBuilder.from_readers(sd.lex_csv, sd.matrix, sd.char_def, sd.unk_def)
od = vo.OracleDictionary(sd.lex_csv, sd.matrix, sd.char_def, sd.unk_def)
utf8, off = synth.make_corpus(sd, 2000, seed=9, log_uniform=(1, 200), unk_frac=0.1)
tok = vb.Tokenizer.new(d2)
tok.set_counting(True)
res = tok.tokenize_batch(utf8=utf8, byte_offsets=off)
tok_off, toks, cnt = od.tokenize_batch(utf8, off, want_counters=True)
assert res.tokens.tobytes() == toks.tobytes() and (tok.last_counters() == cnt).all()
print("synthetic ok", res.n_tokens)
# output stage (k_format_len / k_format_write), connection-id counters, Raw and Dual connector cost functions
for mode in ("mecab", "wakati", "detail"):
    tok.output_mode(mode)
    res = tok.tokenize_batch(utf8=utf8, byte_offsets=off)
    toff, text = res.text()
    eoff, etext = vo.format_batch(od, utf8, off, tok_off, toks, mode)
    assert text == etext and (toff == eoff).all(), mode
    print("output stage", mode, len(text), flush=True)
tok.output_mode(None)
tok.init_connid_counter()
tok.tokenize_batch(utf8=utf8, byte_offsets=off)
print("connid counts", [int(x.sum()) for x in tok.connid_counts()], flush=True)
right, left, cost = synth.make_bigram_files(sd, n_templates=12)
for dual in (False, True):
    dd = vb.SystemDictionaryBuilder.from_readers_with_bigram_info(sd.lex_csv, right, left, cost, sd.char_def, sd.unk_def,
                                                                   dual_connector=dual)
    odd = vo.OracleDictionary(sd.lex_csv, (right, left, cost), sd.char_def, sd.unk_def, dual_connector=dual)
    res = vb.Tokenizer.new(dd).tokenize_batch(utf8=utf8, byte_offsets=off)
    assert res.tokens.tobytes() == odd.tokenize_batch(utf8, off)[1].tobytes()
    print("bigram connector dual =", dual, res.n_tokens, flush=True)
