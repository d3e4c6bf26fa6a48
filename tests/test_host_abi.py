"""CPU-side tests of the product library: the C ABI loads and exports every declared symbol, the
host dictionary model agrees with the oracle, `.dic` round-trips, and the hot path refuses to run
without a GPU (no fallback).  No compute kernels are launched here."""
import re

import numpy as np
import pytest

import vibrato_b200 as vb
from vibrato_b200 import _native, synth
from oracle import vibrato_oracle as vo


def product_dict(golden, user=False):
    r = golden["resources"]
    d = vb.SystemDictionaryBuilder.from_readers(r["lex.csv"], r["matrix.def"], r["char.def"], r["unk.def"])
    if user:
        d = d.reset_user_lexicon_from_reader(r["user.csv"])
    return d


def test_abi_exports_every_declared_symbol():
    hdr = open(_native.HEADER_PATH, encoding="utf-8").read()
    declared = set(re.findall(r"\b(vbt_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    import ctypes
    L = ctypes.CDLL(_native.SO_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} is declared in include/vibrato_b200.h but not exported"
    assert b"sm_100a" in _native.lib().vbt_version()


def test_lexicon_matches_reference_vectors(golden):
    lc = golden["lexicon_cases"]
    d = product_dict(golden)
    for key in ("common_prefix_1", "common_prefix_2"):
        got = d.common_prefix(lc[key]["input"])
        exp = lc[key]["matches"]
        assert [(w, e) for w, e in got] == [(w, e) for w, _, e in exp]
        for (wid, _), (_, eparam, _) in zip(got, exp):
            assert list(d.word_param(wid)) == eparam
    for wid, feat in lc["features"]["items"]:
        assert d.word_feature(wid) == feat
    dup = lc["duplicate_surface"]
    lex = "".join(f"{w},0,0,0,f{i}\n" for i, w in enumerate(dup["words"]))
    d2 = vb.SystemDictionaryBuilder.from_readers(lex, "1 1\n0 0 0", "DEFAULT 0 1 0", "DEFAULT,0,0,100,*")
    assert [list(x) for x in d2.common_prefix(dup["input"])] == dup["matches"]


def test_csv_and_error_cases(golden):
    cs = golden["lexicon_cases"]["csv"]
    mini = ("3 3\n0 0 0", "DEFAULT 0 1 0", "DEFAULT,0,0,100,*")
    for ok in cs["ok"]:
        d = vb.SystemDictionaryBuilder.from_readers(ok["data"], *mini)
        for i, (p, f) in enumerate(zip(ok["params"], ok["features"])):
            assert list(d.word_param(i)) == p and d.word_feature(i) == f
    d = vb.SystemDictionaryBuilder.from_readers(cs["empty_surface"]["data"], *mini)
    assert d.shape()["n_system"] == cs["empty_surface"]["n"]
    kinds = []
    for bad in cs["errors"]:
        with pytest.raises(vb.VibratoError) as ei:
            vb.SystemDictionaryBuilder.from_readers(bad, *mini)
        kinds.append(ei.value.kind)
    assert kinds == ["InvalidFormat", "ParseInt", "ParseInt", "ParseInt"]
    # matrix_connector.rs:233-239 header > u16, builder.rs:151-188 ids outside the matrix
    with pytest.raises(vb.VibratoError):
        vb.SystemDictionaryBuilder.from_readers("a,0,0,1,x\n", "65536 1\n", mini[1], mini[2])
    with pytest.raises(vb.VibratoError):
        vb.SystemDictionaryBuilder.from_readers("a,5,0,1,x\n", *mini)
    with pytest.raises(vb.VibratoError):
        vb.SystemDictionaryBuilder.from_readers("a,0,0,1,x\n", mini[0], mini[1], "DEFAULT,0,7,100,*")
    with pytest.raises(vb.VibratoError):  # char.def without DEFAULT (character.rs:306-311)
        vb.SystemDictionaryBuilder.from_readers("a,0,0,1,x\n", mini[0], "KANJI 0 0 2", mini[2])
    with pytest.raises(vb.VibratoError):  # user lexicon with out-of-range ids (dictionary.rs:218-223)
        product_dict(golden).reset_user_lexicon_from_reader("x,99,0,1,f\n")


def test_ignore_space_requires_space_category():
    d = vb.SystemDictionaryBuilder.from_readers("a,0,0,1,x\n", "1 1\n0 0 0", "DEFAULT 0 1 0", "DEFAULT,0,0,100,*")
    with pytest.raises(vb.VibratoError) as ei:
        vb.Tokenizer.new(d).ignore_space(True)  # tokenizer.rs:44-49
    assert ei.value.kind == "InvalidArgument"
    vb.Tokenizer.new(d).ignore_space(False)


def test_host_dictionary_agrees_with_oracle_on_synthetic():
    sd = synth.make_dictionary("synth-small")
    d = vb.SystemDictionaryBuilder.from_readers(sd.lex_csv, sd.matrix, sd.char_def, sd.unk_def)
    od = vo.OracleDictionary(sd.lex_csv, sd.matrix, sd.char_def, sd.unk_def)
    user = synth.make_user_csv(sd, 300)
    d.reset_user_lexicon_from_reader(user)
    od.set_user_csv(user)
    sh = d.shape()
    assert (sh["num_left"], sh["num_right"]) == (od.num_left, od.num_right)
    assert sh["n_system"] == od.num_words(0) and sh["n_user"] == od.num_words(1) and sh["n_unknown"] == od.num_words(2)
    utf8, off = synth.make_corpus(sd, 300, unk_frac=0.2)
    rng = np.random.default_rng(5)
    for i in range(300):
        text = bytes(utf8[int(off[i]):int(off[i + 1])]).decode()
        st = int(rng.integers(0, max(1, len(text) - 1)))
        for lex in (0, 1):
            assert d.common_prefix(text[st:st + 12], lex) == od.common_prefix(text[st:st + 12], lex)
    for wid in rng.integers(0, sh["n_system"], 200):
        assert d.word_feature(int(wid)) == od.feature(int(wid))
        assert d.word_param(int(wid)) == od.word_param(int(wid))
    for wid in range(sh["n_unknown"]):
        assert d.word_feature((2 << 30) | wid) == od.feature((2 << 30) | wid)
    for wid in range(0, sh["n_user"], 7):
        assert d.word_param((1 << 30) | wid) == od.word_param((1 << 30) | wid)


def test_dic_stream_roundtrip(golden):
    d = product_dict(golden, user=True)
    blob = d.write()
    assert blob.startswith(b"VibratoTokenizer 0.5\n")  # dictionary.rs:27
    d2 = vb.Dictionary.read(blob)
    assert d2.shape() == d.shape()
    assert d2.write() == blob
    for text in ("東京都に行く", "京都東京都", "XX"):
        assert d2.common_prefix(text) == d.common_prefix(text)
        assert d2.common_prefix(text, 1) == d.common_prefix(text, 1)
    assert d2.word_feature((1 << 30) | 0) == "カスタム名詞"
    assert (d2.pack_blob() == d.pack_blob()).all()
    with pytest.raises(vb.VibratoError) as ei:  # dictionary.rs:188-193
        vb.Dictionary.read(b"VibratoTokenizer 0.4\n" + blob[21:])
    assert ei.value.kind == "InvalidArgument"
    with pytest.raises(vb.VibratoError):
        vb.Dictionary.read(blob[: len(blob) // 2])


def test_blob_is_self_consistent(golden):
    d = product_dict(golden)
    b = d.pack_blob()
    assert b.nbytes % 256 == 0 and bytes(b[:8]) == b"VTBLOB04"
    assert int(np.frombuffer(b[8:16].tobytes(), dtype="<u8")[0]) == b.nbytes


def test_hot_path_fails_loudly_without_gpu(golden):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    t = vb.Tokenizer.new(product_dict(golden))
    with pytest.raises(vb.VibratoError) as ei:
        t.tokenize_batch(["東京都"])
    assert ei.value.kind in ("NoDevice", "Cuda")
    w = t.new_worker()
    w.reset_sentence("東京都")
    with pytest.raises(vb.VibratoError):
        w.tokenize()


def test_connector_variant_tags(golden):
    """ConnectorWrapper variants (connector.rs:30-35): a Raw or Dual tag in front of a matrix payload fails to
    decode, unknown tags are decode errors."""
    import struct
    d = product_dict(golden)
    blob = bytearray(d.write())
    # locate the connector tag: it follows `user_lexicon: None` (one zero byte) after the system lexicon
    d_user = product_dict(golden, user=True)
    assert len(d_user.write()) > len(blob)
    marker = struct.pack("<B", 0) + struct.pack("<I", 0) + struct.pack("<Q", 100)  # None, Matrix, Vec len 10*10
    at = bytes(blob).find(marker)
    assert at > 0
    for variant, kind in ((2, "BincodeDecode"), (1, "BincodeDecode"), (7, "BincodeDecode")):
        bad = bytearray(blob)
        bad[at + 1:at + 5] = struct.pack("<I", variant)
        with pytest.raises(vb.VibratoError) as ei:
            vb.Dictionary.read(bytes(bad))
        assert ei.value.kind == kind


def test_rust_shim_binds_only_exported_symbols():
    """Every extern "C" fn the Rust shim declares exists in the library with the header's name."""
    import ctypes
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "rust-shim", "src", "lib.rs"), encoding="utf-8").read()
    names = set(re.findall(r"fn (vbt_[a-z0-9_]+)\(", src))
    assert len(names) >= 12
    L = ctypes.CDLL(_native.SO_PATH)
    hdr = open(_native.HEADER_PATH, encoding="utf-8").read()
    for n in names:
        assert hasattr(L, n) and re.search(r"\b" + n + r"\s*\(", hdr), n


def test_map_connection_ids_matches_reference_vectors_and_oracle(golden):
    """ConnIdMapper::parse vectors (mapper.rs:163-180) through Dictionary::map_connection_ids_from_iter
    (dictionary.rs:245-259), product against oracle, including a user lexicon attached afterwards."""
    lex = "a,1,1,1,x\nb,2,2,1,y\n"
    mat = "5 5\n" + "".join(f"{r} {l} {r * 10 + l}\n" for r in range(5) for l in range(5))
    d = vb.SystemDictionaryBuilder.from_readers(lex, mat, "DEFAULT 0 1 0", "DEFAULT,3,3,100,*")
    od = vo.OracleDictionary(lex, mat, "DEFAULT 0 1 0", "DEFAULT,3,3,100,*")
    d.map_connection_ids_from_iter([2, 3, 4, 1], [2, 3, 4, 1])  # mapper.rs:164-167 -> new_ids [0,4,1,2,3]
    od.map_connection_ids([2, 3, 4, 1], [2, 3, 4, 1])
    assert d.word_param(0) == (4, 4, 1) and d.word_param(1) == (1, 1, 1)
    assert d.word_param((2 << 30) | 0) == (2, 2, 100)
    for r in range(5):
        for l in range(5):
            assert d.conn_cost(r, l) == od.conn_cost(r, l)
    assert d.conn_cost(4, 4) == 11 and d.conn_cost(0, 1) == 2  # old (1,1) and old (right 0, left 2)
    d.reset_user_lexicon_from_reader("c,1,2,5,u\n")  # ids go through the stored mapper (dictionary.rs:215-217)
    od.set_user_csv("c,1,2,5,u\n")
    assert d.word_param((1 << 30) | 0) == od.word_param((1 << 30) | 0) == (4, 1, 5)
    d2 = vb.Dictionary.read(d.write())  # mapper: Some(..) survives the .dic round trip
    d2.reset_user_lexicon_from_reader("c,1,2,5,u\n")
    assert d2.word_param((1 << 30) | 0) == (4, 1, 5)
    for bad, kind in (([2, 3, 0, 1], "InvalidArgument"), ([2, 3, 5, 1], "InvalidArgument"), ([2, 2, 3, 1], "InvalidArgument"),
                      ([1, 2, 3], "InvalidArgument")):
        with pytest.raises(vb.VibratoError) as ei:
            d.map_connection_ids_from_iter(bad, [1, 2, 3, 4])
        assert ei.value.kind == kind
        with pytest.raises(vo.OracleError):
            od.map_connection_ids(bad, [1, 2, 3, 4])
    # ConnIdCounter::compute_probs, mapper.rs:152-161
    lp, rp = vo.compute_connid_probs(np.array([1, 5, 4]), np.array([3, 0, 7]))
    assert lp == [(1, 0.5), (2, 0.4)] and rp == [(2, 0.7), (1, 0.0)]


def test_char_def_and_matrix_def_reference_vectors():
    """character.rs:288-366 and matrix_connector.rs:131-262, product and oracle."""
    mini = ("a,0,0,1,x\n", "1 1\n0 0 0", "DEFAULT,0,0,100,*")
    cd = "DEFAULT 0 1 0\nSPACE 0 1 0\n0x0020 SPACE"
    d = vb.SystemDictionaryBuilder.from_readers(mini[0], mini[1], cd, mini[2])
    od = vo.OracleDictionary(mini[0], mini[1], cd, mini[2])
    for ci in (d.char_info(0x20), od.char_info(0x20)):
        assert ci & 0x3FFFF == 0b10 and (ci >> 18) & 0xFF == 1 and not (ci >> 26) & 1 and (ci >> 27) & 1 and ci >> 28 == 0
    assert d.char_info(0x1F600) == d.char_info(0) == od.char_info(0x1F600)  # character.rs:112-116 fallback to entry 0
    bad_char_defs = ["DEFAULT 0 1 0\n0x0..0xFFFF INVALID", "USER_DEFINED 0 1 0", "DEFAULT 2 1 0", "DEFAULT 0 2 0",
                     "DEFAULT 0 2 -1", "DEFAULT 0 2", "DEFAULT 0 1 0\n0x10000 DEFAULT", "DEFAULT 0 1 0\n0x0..0x10000 DEFAULT",
                     "DEFAULT 0 1 0\n0x0020..0x0019 DEFAULT"]
    for bad in bad_char_defs:
        with pytest.raises(vb.VibratoError):
            vb.SystemDictionaryBuilder.from_readers(mini[0], mini[1], bad, mini[2])
        with pytest.raises(vo.OracleError):
            vo.OracleDictionary(mini[0], mini[1], bad, mini[2])
    vb.SystemDictionaryBuilder.from_readers(mini[0], mini[1], "DEFAULT 0 1 0\n0x0..0xFFFF DEFAULT", mini[2])
    m23 = "2 3\n0 0 0\n0 1 1\n0 2 2\n1 0 -3\n1 1 -4\n1 2 -5"
    d = vb.SystemDictionaryBuilder.from_readers("a,0,0,1,x\n", m23, "DEFAULT 0 1 0", mini[2])
    od = vo.OracleDictionary("a,0,0,1,x\n", m23, "DEFAULT 0 1 0", mini[2])
    for (r, l), c in {(0, 0): 0, (0, 1): 1, (0, 2): 2, (1, 0): -3, (1, 1): -4, (1, 2): -5}.items():
        assert d.conn_cost(r, l) == od.conn_cost(r, l) == c
    bad_matrices = ["2\n0 0 0\n0 1 1\n1 0 -2\n1 1 -3", "2 2 2\n0 0 0\n0 1 1\n1 0 -2\n1 1 -3", "2 2\n0 0 0\n0 1 1\n1 -2\n1 1 -3",
                    "2 2\n0 0 0\n0 1 1\n1 0 1 -2\n1 1 -3", "65536 65536", "2 2\n0 0 0\n0 1 1\n1 2 -2\n1 1 -3",
                    "2 2\n0 0 0\n0 1 1\n2 0 -2\n1 1 -3"]
    for bad in bad_matrices:
        with pytest.raises(vb.VibratoError):
            vb.SystemDictionaryBuilder.from_readers("a,0,0,1,x\n", bad, "DEFAULT 0 1 0", mini[2])
        with pytest.raises(vo.OracleError):
            vo.OracleDictionary("a,0,0,1,x\n", bad, "DEFAULT 0 1 0", mini[2])


RAW_RIGHT = "1\tSURF-SURF:これ,*,SURF-POS:これ,POS-SURF:代名詞,*\n2\tSURF-SURF:テスト,*,SURF-POS:テスト,POS-SURF:名詞,*"
RAW_LEFT = "1\tです,*,助動詞,です,*\n2\tは,*,助詞,は,*"
RAW_COST = "SURF-SURF:これ/は\t-100\nSURF-POS:これ/助詞\t200\nPOS-SURF:代名詞/は\t-300"
SCORER_TRIPLES = [(18, 17, 1), (4, 9, 2), (17, 0, 3), (17, 12, 4), (8, 6, 5), (2, 5, 6), (12, 18, 7), (9, 1, 8), (19, 5, 9),
                  (9, 4, 10), (0, 19, 11), (2, 19, 12), (7, 9, 13), (18, 9, 14), (17, 4, 15), (9, 6, 16), (13, 0, 17),
                  (1, 4, 18), (0, 18, 19), (18, 11, 20)]


def test_raw_connector_reference_vectors():
    """scorer.rs:355-480 (retrieve_cost / accumulate_cost == 100) and raw_connector.rs:467-511
    (from_readers: cost(1,2) == -200; after mapping cost(0,0) == -200), product and oracle."""
    from vibrato_b200.api import scorer_accumulate
    INV = 0x7FFFFFFF
    k1 = [18, 17, 0, INV, 8, 12, 19, INV, INV, 9, 0, 7, 17, 13, 0, INV]
    k2 = [17, 0, 0, INV, 6, 18, 5, INV, INV, 9, 19, 9, 4, 0, 18, INV]
    for acc in (scorer_accumulate, vo.scorer_accumulate):
        assert acc(SCORER_TRIPLES, k1, k2) == 100
        for a, b, e in [(0, 18, 19), (0, 19, 11), (9, 4, 10), (9, 6, 16), (0, 0, 0), (9, 5, 0)]:
            assert acc(SCORER_TRIPLES, [a], [b]) == e
        assert acc([], [], []) == 0
    mini = ("a,1,2,5,x\n", "DEFAULT 0 1 0", "DEFAULT,0,0,100,*")
    d = vb.SystemDictionaryBuilder.from_readers_with_bigram_info(mini[0], RAW_RIGHT, RAW_LEFT, RAW_COST, mini[1], mini[2])
    od = vo.OracleDictionary(mini[0], (RAW_RIGHT, RAW_LEFT, RAW_COST), mini[1], mini[2])
    assert d.shape()["num_left"] == od.num_left == 3 and d.shape()["num_right"] == od.num_right == 3
    assert d.conn_cost(1, 2) == od.conn_cost(1, 2) == -200
    # raw_connector.rs:489-510 maps left (1,2,0) / right (2,0,1) directly; through from_iter id 0 stays, so
    # check the row move with a legal mapping instead: right 1 -> 2, left 2 -> 1
    d.map_connection_ids_from_iter([2, 1], [2, 1])
    od.map_connection_ids([2, 1], [2, 1])
    assert d.conn_cost(2, 1) == od.conn_cost(2, 1) == -200
    d2 = vb.Dictionary.read(d.write())  # Raw variant of the .dic stream round-trips
    assert d2.conn_cost(2, 1) == -200 and d2.write() == d.write()
    with pytest.raises(vb.VibratoError) as ei:  # 2 feature templates: the Dual split needs at least SIMD_SIZE = 8
        vb.SystemDictionaryBuilder.from_readers_with_bigram_info(mini[0], RAW_RIGHT, RAW_LEFT, RAW_COST, mini[1], mini[2],
                                                                 dual_connector=True)
    assert ei.value.kind == "InvalidArgument"
    with pytest.raises(vo.OracleError):
        vo.OracleDictionary(mini[0], (RAW_RIGHT, RAW_LEFT, RAW_COST), mini[1], mini[2], dual_connector=True)
    bad = [(RAW_RIGHT, RAW_LEFT, "SURF-SURF:これは\t100"), (RAW_RIGHT, RAW_LEFT, "SURF-SURF:これ/は100"),
           (RAW_RIGHT, RAW_LEFT, "SURF-SURF:これ/は\tabc"), ("これ,*", RAW_LEFT, RAW_COST), ("2\tこれ", RAW_LEFT, RAW_COST)]
    for r, l, c in bad:  # raw_connector.rs:380-416, 447-464, 214-219
        with pytest.raises(vb.VibratoError):
            vb.SystemDictionaryBuilder.from_readers_with_bigram_info(mini[0], r, l, c, mini[1], mini[2])
        with pytest.raises(vo.OracleError):
            vo.OracleDictionary(mini[0], (r, l, c), mini[1], mini[2])


DUAL_RIGHT = "1\tAB,*,CD,*,EF,*,GH,*,IJ,*,KL,*,MN,*,OP,*,QR,*,ST\n2\tUV,*,WX,*,YZ,*,12,*,34,*,56,*,78,*,90,*,*,*,*"
DUAL_LEFT = "1\tuv,*,wx,*,yz,*,12,*,34,*,56,*,78,*,90,*,*,*,*\n2\tab,*,cd,*,ef,*,gh,*,ij,*,kl,*,mn,*,op,*,qr,*,st"
DUAL_COST = "\n".join(f"{a}\t{c}" for a, c in [
    ("AB/ab", -10), ("CD/cd", 20), ("EF/ef", -30), ("GH/gh", 40), ("IJ/ij", -50), ("KL/kl", 60), ("MN/mn", -70),
    ("OP/op", 80), ("QR/qr", -90), ("ST/st", 100), ("UV/uv", -110), ("WX/wx", 120), ("YZ/yz", -130), ("12/12", 140),
    ("34/34", -150), ("56/56", 160), ("78/78", -170), ("90/90", 180)])


def test_dual_connector_reference_vectors():
    """dual_connector.rs:285-361: from_readers gives cost(1,2) == 50 and cost(2,1) == 40, and the values follow
    the ids through map_connection_ids; product and oracle, plus the Dual variant of the .dic stream."""
    mini = ("a,1,2,5,x\n", "DEFAULT 0 1 0", "DEFAULT,0,0,100,*")
    d = vb.SystemDictionaryBuilder.from_readers_with_bigram_info(mini[0], DUAL_RIGHT, DUAL_LEFT, DUAL_COST, mini[1], mini[2],
                                                                 dual_connector=True)
    od = vo.OracleDictionary(mini[0], (DUAL_RIGHT, DUAL_LEFT, DUAL_COST), mini[1], mini[2], dual_connector=True)
    raw = vb.SystemDictionaryBuilder.from_readers_with_bigram_info(mini[0], DUAL_RIGHT, DUAL_LEFT, DUAL_COST, mini[1], mini[2])
    assert d.shape()["num_left"] == od.num_left == 3 and d.shape()["num_right"] == od.num_right == 3
    for x in (d, od, raw):
        assert x.conn_cost(1, 2) == 50 and x.conn_cost(2, 1) == 40
        assert x.conn_cost(0, 0) == 0 and x.conn_cost(1, 1) == 0 and x.conn_cost(0, 2) == 0
    # dual_connector.rs:325-360 maps left (1,2,0) / right (2,0,1) directly; from_iter keeps id 0, so swap 1 <-> 2
    d.map_connection_ids_from_iter([2, 1], [2, 1])
    od.map_connection_ids([2, 1], [2, 1])
    for x in (d, od):
        assert x.conn_cost(2, 1) == 50 and x.conn_cost(1, 2) == 40 and x.conn_cost(0, 0) == 0
    stream = d.write()
    d2 = vb.Dictionary.read(stream)
    assert d2.conn_cost(2, 1) == 50 and d2.conn_cost(1, 2) == 40 and d2.write() == stream
    assert d2.pack_blob()[:8].tobytes() == b"VTBLOB04"


def test_dual_connector_costs_match_raw_and_oracle_on_synthetic():
    """DualConnector::cost (dual_connector.rs:269-280) is the RawConnector sum stored differently: a reduced
    matrix over all but eight templates plus an 8-lane raw term."""
    sd = synth.make_dictionary("synth-tiny")
    right, left, cost = synth.make_bigram_files(sd, n_templates=12)
    build = vb.SystemDictionaryBuilder.from_readers_with_bigram_info
    d = build(sd.lex_csv, right, left, cost, sd.char_def, sd.unk_def, dual_connector=True)
    raw = build(sd.lex_csv, right, left, cost, sd.char_def, sd.unk_def)
    od = vo.OracleDictionary(sd.lex_csv, (right, left, cost), sd.char_def, sd.unk_def, dual_connector=True)
    assert d.shape()["num_left"] == od.num_left == sd.num_left and d.shape()["num_right"] == od.num_right == sd.num_right
    vals = set()
    for r in range(sd.num_right):
        for l in range(0, sd.num_left, 3):
            c = d.conn_cost(r, l)
            assert c == od.conn_cost(r, l) == raw.conn_cost(r, l)
            vals.add(c)
    assert len(vals) > 50 and d.conn_cost(0, 0) == 0
    lmap = list(range(sd.num_left - 1, 0, -1))
    rmap = list(range(sd.num_right - 1, 0, -1))
    d.map_connection_ids_from_iter(lmap, rmap)
    od.map_connection_ids(lmap, rmap)
    raw.map_connection_ids_from_iter(lmap, rmap)
    for r in range(0, sd.num_right, 2):
        for l in range(0, sd.num_left, 5):
            assert d.conn_cost(r, l) == od.conn_cost(r, l) == raw.conn_cost(r, l)
    d2 = vb.Dictionary.read(d.write())
    assert all(d2.conn_cost(r, 1) == d.conn_cost(r, 1) for r in range(sd.num_right))


def test_raw_connector_costs_match_oracle_on_synthetic():
    sd = synth.make_dictionary("synth-tiny")
    right, left, cost = synth.make_bigram_files(sd)
    d = vb.SystemDictionaryBuilder.from_readers_with_bigram_info(sd.lex_csv, right, left, cost, sd.char_def, sd.unk_def)
    od = vo.OracleDictionary(sd.lex_csv, (right, left, cost), sd.char_def, sd.unk_def)
    assert d.shape()["num_left"] == od.num_left == sd.num_left and d.shape()["num_right"] == od.num_right == sd.num_right
    vals = set()
    for r in range(sd.num_right):
        for l in range(0, sd.num_left, 3):
            c = d.conn_cost(r, l)
            assert c == od.conn_cost(r, l)
            vals.add(c)
    assert len(vals) > 50 and d.conn_cost(0, 0) == 0
    assert (d.pack_blob()[:8].tobytes() == b"VTBLOB04")


def test_dic_reader_survives_corrupted_streams(golden):
    """`.dic` files are untrusted input: truncated or byte-flipped streams must end in a VibratoError or in a
    dictionary whose device image still packs or is refused — never in a crash.  (A flipped trie node once sent
    Trie::enumerate out of bounds: kept as the first case.)"""
    d = product_dict(golden, user=True)
    good = bytes(d.write())
    rng = np.random.default_rng(20260923)
    cases = [[(149563, 79)]] + [[(int(rng.integers(21, len(good))), int(rng.integers(0, 256)))
                                 for _ in range(int(rng.integers(1, 4)))] for _ in range(400)]
    accepted = refused = 0
    for muts in cases:
        b = bytearray(good)
        for pos, v in muts:
            if pos < len(b):
                b[pos] = v
        try:
            dd = vb.Dictionary.read(bytes(b))
            accepted += 1
            try:
                dd.pack_blob()
            except vb.VibratoError:
                pass
        except vb.VibratoError:
            refused += 1
    assert accepted > 50 and refused > 10
    for cut in list(range(0, 64)) + [int(x) for x in rng.integers(64, len(good), 100)]:
        with pytest.raises(vb.VibratoError):
            vb.Dictionary.read(good[:cut])


def test_connection_id_mapping_preserves_costs_for_every_connector():
    """Dictionary::map_connection_ids_from_iter (dictionary.rs:245-259) only renames ids: for the Matrix, Raw and Dual
    connectors cost(new_right, new_left) == cost(old_right, old_left), also after a trip through the .dic stream;
    a second mapping composes with the first."""
    sd = synth.make_dictionary("synth-tiny")
    right, left, cost = synth.make_bigram_files(sd, n_templates=12)
    build = vb.SystemDictionaryBuilder
    dicts_ = [build.from_readers(sd.lex_csv, sd.matrix, sd.char_def, sd.unk_def),
              build.from_readers_with_bigram_info(sd.lex_csv, right, left, cost, sd.char_def, sd.unk_def),
              build.from_readers_with_bigram_info(sd.lex_csv, right, left, cost, sd.char_def, sd.unk_def, dual_connector=True)]
    rng = np.random.default_rng(11)
    for d in dicts_:
        nl, nr = d.shape()["num_left"], d.shape()["num_right"]
        pairs = [(int(rng.integers(0, nr)), int(rng.integers(0, nl))) for _ in range(300)] + [(0, 0), (0, nl - 1), (nr - 1, 0)]
        before = [d.conn_cost(r, l) for r, l in pairs]
        new_l, new_r = np.arange(nl), np.arange(nr)  # current id of each original id
        for _ in range(2):
            lmap = rng.permutation(np.arange(1, nl))
            rmap = rng.permutation(np.arange(1, nr))
            d.map_connection_ids_from_iter([int(x) for x in lmap], [int(x) for x in rmap])
            step_l, step_r = np.zeros(nl, dtype=np.int64), np.zeros(nr, dtype=np.int64)
            step_l[lmap] = np.arange(1, nl)
            step_r[rmap] = np.arange(1, nr)
            new_l, new_r = step_l[new_l], step_r[new_r]
            assert [d.conn_cost(int(new_r[r]), int(new_l[l])) for r, l in pairs] == before
        d2 = vb.Dictionary.read(d.write())
        assert [d2.conn_cost(int(new_r[r]), int(new_l[l])) for r, l in pairs] == before


def test_host_parsers_agree_with_oracle_on_mutated_sources(golden):
    """Differential check of the two independent restatements of the MeCab-source parsers (lex.csv / matrix.def /
    char.def / unk.def / user.csv): on mutated copies of the reference's own fixture they must accept or refuse
    together, name the same VibratoError variant, and answer lookups identically."""
    r = golden["resources"]
    names = ["lex.csv", "matrix.def", "char.def", "unk.def", "user.csv"]
    src = [r[k].encode() for k in names]
    rng = np.random.default_rng(20260924)
    bits = [b",", b'"', b"\n", b"\r\n", b"\t", b" ", b"0", b"9", b"-", b"+", b"/", b"*", b"#", b"x", b"0x", b"..",
            "あ".encode(), b" ,", b"1 ", b"-0", b"65535", b"65536", b"32768", b"-32769"]

    def mutate(b):
        b = bytearray(b)
        for _ in range(int(rng.integers(1, 4))):
            pos = int(rng.integers(0, len(b)))
            kind = int(rng.integers(0, 4))
            tok = bits[int(rng.integers(0, len(bits)))]
            if kind == 0:
                b[pos:pos + 1] = tok
            elif kind == 1:
                del b[pos:pos + int(rng.integers(1, 6))]
            elif kind == 2:
                b[pos:pos] = tok
            else:  # duplicate the line
                s0 = b.rfind(b"\n", 0, pos) + 1
                e0 = b.find(b"\n", pos)
                e0 = len(b) if e0 < 0 else e0 + 1
                b[s0:s0] = b[s0:e0]
        return bytes(b)

    accepted = refused = 0
    while accepted + refused < 160:
        t = list(src)
        which = int(rng.integers(0, 5))
        t[which] = mutate(t[which])
        try:
            t[which].decode("utf-8")
        except UnicodeDecodeError:
            continue
        pd = od = perr = oerr = None
        try:
            pd = vb.SystemDictionaryBuilder.from_readers(t[0], t[1], t[2], t[3])
            pd.reset_user_lexicon_from_reader(t[4])
        except vb.VibratoError as e:
            pd, perr = None, e
        try:
            od = vo.OracleDictionary(t[0], t[1], t[2], t[3])
            od.set_user_csv(t[4])
        except vo.OracleError as e:
            od, oerr = None, e
        assert (pd is None) == (od is None), (names[which], perr, oerr)
        if pd is None:
            refused += 1
            assert perr.kind == str(oerr).split("(")[0], (names[which], perr, oerr)
            continue
        accepted += 1
        sh = pd.shape()
        assert (sh["num_left"], sh["num_right"]) == (od.num_left, od.num_right)
        assert all(pd.char_info(cp) == od.char_info(cp) for cp in (0x3042, 0x4EAC, 0x41, 0x20, 0x30, 0x10000, 0xFFFF, 0x3000))
        assert all(pd.conn_cost(a, b) == od.conn_cost(a, b) for a in range(0, sh["num_right"], 3) for b in range(0, sh["num_left"], 3))
        for lex, key in ((0, "n_system"), (1, "n_user")):
            assert sh[key] == od.num_words(lex)
            for wid in range(sh[key]):
                w = (lex << 30) | wid
                assert pd.word_feature(w) == od.feature(w) and tuple(pd.word_param(w)) == tuple(od.word_param(w))
            for text in ("京都東京都京都", "東京都", "自然言語処理", "kampersanda", "本とカレー"):
                assert pd.common_prefix(text, lex) == od.common_prefix(text, lex)
    assert accepted > 20 and refused > 20


def test_bigram_builders_agree_with_oracle_on_mutated_sources():
    """Same differential check for from_readers_with_bigram_info (Raw and Dual connectors) on mutated bigram.* files."""
    sd = synth.make_dictionary("synth-tiny")
    right, left, cost = synth.make_bigram_files(sd, n_templates=10)
    src = [x.encode() if isinstance(x, str) else bytes(x) for x in (right, left, cost)]
    rng = np.random.default_rng(20260925)
    bits = [b",", b'"', b"\n", b"\r\n", b"\t", b" ", b"0", b"9", b"-", b"+", b"/", b"*", b"#", b"x", "あ".encode(),
            b"\t\t", b"//", b'""', b"2147483648", b"-2147483649"]
    accepted = refused = 0
    while accepted + refused < 80:
        t = [bytearray(x) for x in src]
        which = int(rng.integers(0, 3))
        b = t[which]
        for _ in range(int(rng.integers(1, 4))):
            pos = int(rng.integers(0, len(b)))
            kind = int(rng.integers(0, 3))
            tok = bits[int(rng.integers(0, len(bits)))]
            if kind == 0:
                b[pos:pos + 1] = tok
            elif kind == 1:
                del b[pos:pos + int(rng.integers(1, 6))]
            else:
                b[pos:pos] = tok
        try:
            b.decode("utf-8")
        except UnicodeDecodeError:
            continue
        t = [bytes(x) for x in t]
        dual = bool(rng.integers(0, 2))
        pd = od = perr = oerr = None
        try:
            pd = vb.SystemDictionaryBuilder.from_readers_with_bigram_info(sd.lex_csv, t[0], t[1], t[2], sd.char_def, sd.unk_def,
                                                                           dual_connector=dual)
        except vb.VibratoError as e:
            perr = e
        try:
            od = vo.OracleDictionary(sd.lex_csv, (t[0], t[1], t[2]), sd.char_def, sd.unk_def, dual_connector=dual)
        except vo.OracleError as e:
            oerr = e
        assert (pd is None) == (od is None), (which, dual, perr, oerr)
        if pd is None:
            refused += 1
            assert perr.kind == str(oerr).split("(")[0], (which, dual, perr, oerr)
            continue
        accepted += 1
        sh = pd.shape()
        assert (sh["num_left"], sh["num_right"]) == (od.num_left, od.num_right)
        assert all(pd.conn_cost(a, c) == od.conn_cost(a, c) for a in range(0, sh["num_right"], 2) for c in range(0, sh["num_left"], 2))
    assert accepted > 10 and refused > 10


def test_zstd_compressed_dictionary_file(golden, tmp_path):
    """`tokenize -i system.dic.zst` (tokenize/src/main.rs:59-60): zstd frame -> Dictionary::read.  The frame is made with
    the same runtime libzstd the loader uses; corrupted or truncated frames and missing files are errors, not crashes."""
    import ctypes as C
    try:
        z = C.CDLL("libzstd.so.1")
    except OSError:
        pytest.skip("libzstd.so.1 not present")
    z.ZSTD_compressBound.restype = C.c_size_t
    z.ZSTD_compressBound.argtypes = [C.c_size_t]
    z.ZSTD_compress.restype = C.c_size_t
    z.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
    d = product_dict(golden, user=True)
    stream = bytes(d.write())
    buf = C.create_string_buffer(z.ZSTD_compressBound(len(stream)))
    n = z.ZSTD_compress(buf, len(buf), stream, len(stream), 3)
    comp = buf.raw[:n]
    path = tmp_path / "system.dic.zst"
    path.write_bytes(comp)
    d2 = vb.Dictionary.from_zstd_file(path)
    assert d2.write() == stream
    rng = np.random.default_rng(3)
    for k in range(60):
        b = bytearray(comp)
        b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
        if k % 5 == 0:
            b = b[:int(rng.integers(0, len(b)))]
        path.write_bytes(bytes(b))
        try:
            vb.Dictionary.from_zstd_file(path)
        except vb.VibratoError:
            pass
    for missing in (tmp_path / "nope.zst", tmp_path):
        with pytest.raises(vb.VibratoError) as ei:
            vb.Dictionary.from_zstd_file(missing)
        assert ei.value.kind == "StdIo"


def test_clis_refuse_to_run_without_a_gpu(golden, tmp_path):
    """The look-alike CLIs load the dictionary on the host and must then stop with an error when no CUDA device
    exists — there is no CPU tokenisation path to fall back to.  (Skipped on a GPU box.)"""
    import os
    import subprocess
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for k, v in golden["resources"].items():
        (tmp_path / k).write_text(v, encoding="utf-8", newline="")
    (tmp_path / "gold.txt").write_text("京都\tX\nEOS\n", encoding="utf-8")
    runs = [(["tokenize", "-i", str(tmp_path)], "京都\n"), (["benchmark", "-i", str(tmp_path)], "京都\n"),
            (["evaluate", "-i", str(tmp_path), "-t", str(tmp_path / "gold.txt")], "")]
    for argv, stdin in runs:
        exe = os.path.join(root, "vibrato_b200", "bin", argv[0])
        if not os.path.exists(exe):
            pytest.skip("CLIs not built")
        p = subprocess.run([exe] + argv[1:], input=stdin.encode(), capture_output=True, timeout=120)
        assert p.returncode != 0 and p.stdout == b"", argv[0]
        assert b"Error: no CUDA device available" in p.stderr, p.stderr


# ---- `.dic` layout against a second derivation, and properties that need no second implementation ----------------

def test_dictionary_read_accepts_an_independently_encoded_stream(golden):
    """Dictionary::read (dictionary.rs:173-197) against tests/dic_format.py — a `.dic` writer derived from the
    reference's struct definitions, sharing nothing with host_dict.cpp.  The loaded
    dictionary must equal the one the product builds from the same sources, and the product's own writer must
    produce the same bytes outside the trie blobs (two double-array builders may place nodes differently)."""
    import dic_format as df
    r = golden["resources"]
    for user in (False, True):
        stream = df.dictionary_bytes(r["lex.csv"], r["matrix.def"], r["char.def"], r["unk.def"],
                                     user_csv=r["user.csv"] if user else None)
        d_py = vb.Dictionary.read(stream)
        d = product_dict(golden, user=user)
        assert d_py.shape() == d.shape()
        sh = d.shape()
        for wid in range(sh["n_system"]):
            assert d_py.word_feature(wid) == d.word_feature(wid)
            assert d_py.word_param(wid) == d.word_param(wid)
        for wid in range(sh["n_user"]):
            assert d_py.word_param((1 << 30) | wid) == d.word_param((1 << 30) | wid)
            assert d_py.word_feature((1 << 30) | wid) == d.word_feature((1 << 30) | wid)
        for wid in range(sh["n_unknown"]):
            assert d_py.word_feature((2 << 30) | wid) == d.word_feature((2 << 30) | wid)
            assert d_py.word_param((2 << 30) | wid) == d.word_param((2 << 30) | wid)
        for right in range(sh["num_right"]):
            for left in range(sh["num_left"]):
                assert d_py.conn_cost(right, left) == d.conn_cost(right, left)
        for cp in list(range(0, 0x250)) + [0x3042, 0x30A2, 0x4E00, 0x4E8C, 0x9FA5, 0xFF10, 0xFFFF, 0x1F600]:
            assert d_py.char_info(cp) == d.char_info(cp)
        for text in ("東京都に行く", "京都東京都京都", "自然言語処理", "XX", "本とカレーの街神保町へようこそ。"):
            for st in range(len(text)):
                for lex in ((0, 1) if user else (0,)):
                    assert d_py.common_prefix(text[st:], lex) == d.common_prefix(text[st:], lex)
        for lex in ((0, 1) if user else (0,)):
            a = d_py.audit(lex)
            assert a == d.audit(lex)
            assert a["words"] == a["listed"] and a["keys_not_found"] == 0 and a["words_unlisted_or_twice"] == 0
        assert df.strip_trie_blobs(d.write(), user) == df.strip_trie_blobs(stream, user)


def test_trie_properties_without_a_second_implementation():
    """Properties of the word map that need no oracle: every inserted key is found by its own lookup, the
    common-prefix result equals a brute-force scan of the key list, and write(read(x)) == x."""
    rng = np.random.default_rng(5)
    alphabet = [chr(c) for c in list(range(0x3042, 0x3060)) + list(range(0x4E00, 0x4E20)) + [0x61, 0x62, 0x1F600]]
    keys = {}
    while len(keys) < 600:
        k = "".join(rng.choice(alphabet, size=int(rng.integers(1, 7))))
        keys.setdefault(k, []).append(len(keys))
    rows, word_of = [], {}
    for k in keys:
        for j in range(1 + int(rng.integers(0, 3))):  # homographs
            word_of.setdefault(k, []).append(len(rows))
            rows.append(f"{k},{int(rng.integers(0, 4))},{int(rng.integers(0, 4))},{int(rng.integers(-500, 5000))},f{len(rows)}")
    order = rng.permutation(len(rows))  # CSV order is arbitrary: word ids follow it
    lex = "\n".join(rows[i] for i in order) + "\n"
    wid_of_row = {int(r): i for i, r in enumerate(order)}
    matrix = "4 4\n" + "".join(f"{r} {l} {r * 7 - l * 3}\n" for r in range(4) for l in range(4))
    d = vb.SystemDictionaryBuilder.from_readers(lex, matrix, "DEFAULT 0 1 0\n", "DEFAULT,0,0,100,*\n")
    a = d.audit(0)
    assert a["keys"] == len(keys) and a["words"] == len(rows) == a["listed"]
    assert a["keys_not_found"] == 0 and a["words_unlisted_or_twice"] == 0
    assert a["longest_key"] == max(len(k) for k in keys)
    texts = list(keys)[:200] + ["".join(rng.choice(alphabet, size=9)) for _ in range(200)]
    for t in texts:
        got = d.common_prefix(t)
        want = []
        for end in range(1, len(t) + 1):  # brute force: every prefix that is a key, ascending length
            if t[:end] in word_of:
                want += [(wid_of_row[r], end) for r in sorted(word_of[t[:end]], key=lambda r: wid_of_row[r])]
        assert got == want, t
    stream = d.write()
    assert vb.Dictionary.read(stream).write() == stream
    # remapping connection ids twice = remapping once with the composed permutation
    d1 = vb.SystemDictionaryBuilder.from_readers(lex, matrix, "DEFAULT 0 1 0\n", "DEFAULT,0,0,100,*\n")
    d2 = vb.SystemDictionaryBuilder.from_readers(lex, matrix, "DEFAULT 0 1 0\n", "DEFAULT,0,0,100,*\n")
    p, q = [2, 3, 1], [3, 1, 2]  # new id order of ids 1..3 (id 0 stays)
    d1.map_connection_ids_from_iter(p, p)
    d1.map_connection_ids_from_iter(q, q)
    # composed: the id that ends up at new position i after both steps
    first = {old: new for new, old in enumerate(p, start=1)}
    second = {old: new for new, old in enumerate(q, start=1)}
    comp_new = {old: second[first[old]] for old in (1, 2, 3)}
    comp = [old for old, _ in sorted(comp_new.items(), key=lambda kv: kv[1])]
    d2.map_connection_ids_from_iter(comp, comp)
    for wid in range(len(rows)):
        assert d1.word_param(wid) == d2.word_param(wid)
    for r in range(4):
        for l in range(4):
            assert d1.conn_cost(r, l) == d2.conn_cost(r, l)


def test_multi_device_tokenizer_fails_loudly_without_gpu(golden):
    """vbt_tokenizer_new_multi has no CPU path either; argument errors are reported before any device is touched."""
    import torch
    d = product_dict(golden)
    with pytest.raises(vb.VibratoError) as ei:  # an empty device list is an argument error everywhere
        vb.Tokenizer.new(d, devices=[]).handle()
    assert ei.value.kind == "InvalidArgument"
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(vb.VibratoError) as ei:
        vb.Tokenizer.new(d, devices=[0, 1]).tokenize_batch(["東京都"])
    assert ei.value.kind in ("NoDevice", "Cuda")
    assert _native.lib().vbt_pin_thread_to_device(0) == 0  # affinity helper: a no-op where sysfs has nothing to say


def test_dict_audit_reports_the_fixture_lexicons(golden):
    """vbt_dict_audit on the fixture dictionary: every key found again, every word listed exactly once; bad arguments."""
    d = product_dict(golden, user=True)
    sh = d.shape()
    a = d.audit(0)
    assert a["words"] == sh["n_system"] == a["listed"] and a["keys_not_found"] == 0 and a["words_unlisted_or_twice"] == 0
    assert 0 < a["keys"] <= a["words"] and a["longest_key"] >= 1
    u = d.audit(1)
    assert u["words"] == sh["n_user"] == u["listed"] and u["keys_not_found"] == 0
    with pytest.raises(vb.VibratoError):
        product_dict(golden).audit(1)  # no user lexicon attached
    with pytest.raises(vb.VibratoError):
        d.audit(2)


def test_shard_rule_is_the_same_in_python_and_in_the_library_description():
    """The byte-balanced split used by bench.py (vibrato_b200.distributed.shard_by_bytes) follows the library's rule
    (multi_engine.cu split_by_bytes): cut i at the first sentence whose start offset reaches first + total / n * i."""
    from vibrato_b200 import distributed as vd
    rng = np.random.default_rng(3)
    for _ in range(50):
        n = int(rng.integers(0, 200))
        off = np.concatenate([[int(rng.integers(0, 50))], rng.integers(0, 40, n)]).cumsum().astype(np.uint64)
        for world in (1, 2, 3, 8):
            sh = vd.shard_by_bytes(off, world)
            assert sh[0][0] == 0 and sh[-1][1] == n and all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
            total = int(off[-1] - off[0])
            for i in range(1, world):
                target = int(off[0]) + total // world * i
                cut = sh[i][0]
                want = max(sh[i - 1][0], min(int(np.searchsorted(off, target, side="left")), n))
                assert cut == want


def test_compact_token_expansion_rebuilds_character_ranges():
    """expand_compact_tokens: the character range of a token = characters of the sentence in front of its byte range
    (the inverse of Sentence::compile's c2b table, sentence.rs:40-46) — checked against oracle tokens, which carry both."""
    from oracle import vibrato_oracle as vo
    from vibrato_b200 import synth
    sd = synth.make_dictionary("synth-tiny")
    od = vo.OracleDictionary(sd.lex_csv, sd.matrix, sd.char_def, sd.unk_def)
    utf8, off = synth.make_corpus(sd, 300, seed=4, log_uniform=(1, 120), unk_frac=0.2, space_frac=0.05, astral_frac=0.02)
    tok_off, toks, _ = od.tokenize_batch(utf8, off, ignore_space=True)
    compact = np.zeros(len(toks), dtype=vb.COMPACT_TOKEN_DTYPE)
    for name in vb.COMPACT_TOKEN_DTYPE.names:
        compact[name] = toks[name]
    full = vb.expand_compact_tokens(compact, tok_off, utf8, off)
    for name in vb.TOKEN_DTYPE.names:
        np.testing.assert_array_equal(full[name], toks[name], err_msg=name)
    assert len(vb.expand_compact_tokens(compact[:0], np.zeros(1, dtype=np.uint64), utf8[:0], np.zeros(1, dtype=np.uint64))) == 0
