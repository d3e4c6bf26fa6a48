"""Pins the oracle (oracle/vibrato_oracle.c) against every golden vector the reference's own unit
tests hold for the tokenisation path (tests/golden/vibrato_fixture.json, generated from
vibrato/src/tests/* by tests/golden/make_golden.py)."""
import numpy as np
import pytest

from oracle import vibrato_oracle as vo


def fixture_dict(golden, user=False):
    r = golden["resources"]
    d = vo.OracleDictionary(r["lex.csv"], r["matrix.def"], r["char.def"], r["unk.def"])
    if user:
        d.set_user_csv(r["user.csv"])
    return d


def check_tokens(got, exp_tokens):
    assert len(got) == len(exp_tokens)
    for g, e in zip(got, exp_tokens):
        assert g["surface"] == e["surface"]
        assert g["range_char"] == e["range_char"]
        assert g["range_byte"] == e["range_byte"]
        if "feature" in e:
            assert g["feature"] == e["feature"]
        if "total_cost" in e:
            assert g["total_cost"] == e["total_cost"]


def test_tokenizer_cases(golden):
    # vibrato/src/tests/tokenizer.rs (15 tests)
    for case in golden["tokenizer_cases"]:
        d = fixture_dict(golden, case["user"])
        w = d.worker(case["ignore_space"], case["max_grouping_len"])
        got = w.tokenize(case["input"])
        if "tokens" in case:
            check_tokens(got, case["tokens"])
        else:
            assert len(got) == case["num_tokens"], case["name"]


def test_repeat(golden):
    d = fixture_dict(golden)
    w = d.worker()
    for text, n in golden["repeat_case"]["sequence"]:
        assert len(w.tokenize(text)) == n


def test_mini_cases(golden):
    # vibrato/src/tokenizer.rs:208-361
    for case in golden["mini_cases"]:
        d = vo.OracleDictionary(case["lex"], case["matrix"], case["char"], case["unk"])
        check_tokens(d.worker().tokenize(case["input"]), case["tokens"])


def test_lexicon_cases(golden):
    lc = golden["lexicon_cases"]
    d = fixture_dict(golden)
    for key in ("common_prefix_1", "common_prefix_2"):
        got = d.common_prefix(lc[key]["input"])
        exp = lc[key]["matches"]
        assert len(got) == len(exp)
        for (wid, end), (ewid, eparam, eend) in zip(got, exp):
            assert (wid, end) == (ewid, eend)
            assert list(d.word_param(wid)) == eparam
    for wid, feat in lc["features"]["items"]:
        assert d.feature(wid) == feat
    dup = lc["duplicate_surface"]
    lex = "".join(f"{w},0,0,0,f{i}\n" for i, w in enumerate(dup["words"]))
    d2 = vo.OracleDictionary(lex, "1 1\n0 0 0", "DEFAULT 0 1 0", "DEFAULT,0,0,100,*")
    assert [list(x) for x in d2.common_prefix(dup["input"])] == dup["matches"]


def test_csv_cases(golden):
    cs = golden["lexicon_cases"]["csv"]
    mini = ("3 3\n0 0 0", "DEFAULT 0 1 0", "DEFAULT,0,0,100,*")
    for ok in cs["ok"]:
        d = vo.OracleDictionary(ok["data"], *mini)
        for i, (p, f) in enumerate(zip(ok["params"], ok["features"])):
            assert list(d.word_param(i)) == p
            assert d.feature(i) == f
    d = vo.OracleDictionary(cs["empty_surface"]["data"], *mini)
    assert d.num_words(0) == cs["empty_surface"]["n"]
    for bad in cs["errors"]:
        with pytest.raises(vo.OracleError):
            vo.OracleDictionary(bad, *mini)


def test_matrix(golden):
    d = fixture_dict(golden)
    m = golden["matrix_cases"]
    assert d.num_left == m["num_left"] and d.num_right == m["num_right"]
    for r, l, c in m["cost"]:
        assert d.conn_cost(r, l) == c


def test_ignore_space_needs_space_category():
    d = vo.OracleDictionary("a,0,0,1,x\n", "1 1\n0 0 0", "DEFAULT 0 1 0", "DEFAULT,0,0,100,*")
    with pytest.raises(vo.OracleError):
        d.worker(ignore_space=True)  # tokenizer.rs:44-49


def test_counters_match_survey(golden):
    # SURVEY.md §8(d) worked values (an independent restatement's numbers)
    for text, user, ign, mgl, walks, U, Cc, M, T, P, W, E, N, K, balg in golden["counter_cases"]:
        d = fixture_dict(golden, user)
        w = d.worker(ign, mgl)
        cnt = np.zeros(vo.NUM_COUNTERS, dtype=np.uint64)
        w.tokenize(text, counters=cnt)
        assert list(map(int, cnt)) == [U, Cc, M, T, P, W, E, N, K, walks], text
        assert int((cnt * vo.B_ALG_WEIGHTS).sum()) == balg


def test_batch_matches_single(golden):
    d = fixture_dict(golden, True)
    texts = [c["input"] for c in golden["tokenizer_cases"]] + ["", "東京 都 ", "X" * 40, "1234京都"]
    blobs = [t.encode() for t in texts]
    off = np.zeros(len(blobs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(b) for b in blobs])
    utf8 = b"".join(blobs)
    for threads in (1, 3):
        tok_off, toks, cnt = d.tokenize_batch(utf8, off, n_threads=threads, want_counters=True)
        w = d.worker()
        for i, t in enumerate(texts):
            single = w.tokenize(t)
            seg = toks[int(tok_off[i]):int(tok_off[i + 1])]
            assert len(single) == len(seg)
            for a, b in zip(single, seg):
                assert a["range_char"] == [int(b["start_char"]), int(b["end_char"])]
                assert a["word_idx"] == int(b["word_idx"]) and a["total_cost"] == int(b["total_cost"])
        assert int(cnt[8]) == len(toks)


def test_utf8_validation():
    assert vo.utf8_valid("東京🗼a".encode())
    for bad in (b"\xff", b"\xc0\x80", b"\xe0\x80\x80", b"\xed\xa0\x80", b"\xf4\x90\x80\x80", b"\xe3\x81"):
        assert not vo.utf8_valid(bad)


def test_dead_end_lattice_yields_no_tokens(golden):
    # fixture unk.def has no KATAKANA entry: "ア" produces no node, EOS has no predecessor.  The
    # reference panics (lattice.rs:163, index u16::MAX); oracle and product define "no tokens".
    d = fixture_dict(golden)
    assert d.worker().tokenize("ア") == []
    assert d.worker().tokenize("東京ア") == []
    assert len(d.worker().tokenize("東京")) == 1
