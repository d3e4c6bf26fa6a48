"""BASELINE.json configs 2-5 at their FULL sizes on the GPU, compared token-for-token with the oracle
(the oracle runs on all host cores; it finishes the 1 M-sentence batch in seconds on the GPU box).
Dictionaries are the seeded synthetic stand-ins of vibrato_b200/synth.py (no real ipadic / unidic here)."""
import os

import numpy as np
import pytest

import vibrato_b200 as vb
from vibrato_b200 import synth
from oracle import vibrato_oracle as vo

pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 8
_cache = {}


def pair(name, user_rows=0):
    key = (name, user_rows)
    if key not in _cache:
        sd = synth.make_dictionary(name)
        d = vb.SystemDictionaryBuilder.from_readers(sd.lex_csv, sd.matrix, sd.char_def, sd.unk_def)
        od = vo.OracleDictionary(sd.lex_csv, sd.matrix, sd.char_def, sd.unk_def)
        ucsv = None
        if user_rows:
            ucsv = synth.make_user_csv(sd, user_rows)
            d.reset_user_lexicon_from_reader(ucsv)
            od.set_user_csv(ucsv)
        _cache[key] = (sd, d, od, ucsv)
    return _cache[key]


def run_and_compare(d, od, utf8, off, ignore_space=False, max_grouping=0):
    tok = vb.Tokenizer.new(d).ignore_space(ignore_space).max_grouping_len(max_grouping)
    res = tok.tokenize_batch(utf8=utf8, byte_offsets=off)
    tok_off, toks, _ = od.tokenize_batch(utf8, off, ignore_space, max_grouping, n_threads=THREADS)
    np.testing.assert_array_equal(res.tok_offsets, tok_off)
    assert res.tokens.tobytes() == toks.tobytes()
    # size-independent properties: tokens tile each sentence in order, costs are cumulative
    t = res.tokens
    assert (t["end_char"] > t["start_char"]).all() and (t["end_byte"] > t["start_byte"]).all()
    first = np.zeros(len(t), dtype=bool)
    first[res.tok_offsets[:-1][res.tok_offsets[:-1] < len(t)].astype(np.int64)] = True
    same_sent = ~first[1:]
    assert (t["start_char"][1:][same_sent] >= t["end_char"][:-1][same_sent]).all()
    return res


def test_config2_ipadic_100k():
    sd, d, od, _ = pair("synth-ipadic")
    utf8, off = synth.make_corpus(sd, 100000, seed=20260923 + 1)
    run_and_compare(d, od, utf8, off)


def test_config4_user_dictionary_mixed_lengths():
    sd, d, od, ucsv = pair("synth-ipadic", 1000)
    utf8, off = synth.make_corpus(sd, 100000, seed=20260923 + 3, log_uniform=(8, 256), unk_frac=0.15,
                                  space_frac=0.02, user_csv=ucsv, user_frac=0.05)
    res = run_and_compare(d, od, utf8, off, ignore_space=True, max_grouping=24)
    lex = res.tokens["word_idx"] >> 30
    assert (lex == 1).sum() > 1000 and (lex == 2).sum() > 10000  # user and unknown words both occur


def test_config3_unidic_1m_and_config5_512_chars():
    sd, d, od, _ = pair("synth-unidic")
    utf8, off = synth.make_corpus(sd, 1000000, seed=20260923 + 2)
    run_and_compare(d, od, utf8, off)
    utf8, off = synth.make_corpus(sd, 10000, seed=20260923 + 4, fixed_len=512)
    run_and_compare(d, od, utf8, off)
