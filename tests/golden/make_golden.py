#!/usr/bin/env python3
"""Generates tests/golden/vibrato_fixture.json from the reference's own unit-test fixtures.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The JSON bundles
  * the MeCab-format fixture dictionary the reference's golden tests are built from
    (vibrato/src/tests/resources/{lex.csv,matrix.def,char.def,unk.def,user.csv} — test DATA, loaded
    by `include_str!` at vibrato/src/tests/tokenizer.rs:4-8), and
  * the expected values asserted by those tests, transcribed below with the file:line of each
    assertion, so the parity tests can run where /root/reference is absent.
Nothing here is reference *source code*.
"""
import json
import os

REF = "/root/reference/vibrato/src/tests/resources"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vibrato_fixture.json")

F_KYOTO = "京都,名詞,固有名詞,地名,一般,*,*,キョウト,京都,*,A,*,*,*,1/5"
F_TOKYOTO = "東京都,名詞,固有名詞,地名,一般,*,*,トウキョウト,東京都,*,B,5/9,*,5/9,*"
F_TOKYO = "東京,名詞,固有名詞,地名,一般,*,*,トウキョウ,東京,*,A,*,*,*,*"
F_TO = "都,名詞,普通名詞,一般,*,*,*,ト,都,*,A,*,*,*,*"
F_ALPHA = "名詞,普通名詞,一般,*,*,*"


def tok(surface, rc, rb, feature=None, total_cost=None):
    d = {"surface": surface, "range_char": list(rc), "range_byte": list(rb)}
    if feature is not None:
        d["feature"] = feature
    if total_cost is not None:
        d["total_cost"] = total_cost
    return d


# vibrato/src/tests/tokenizer.rs — dictionary = the four fixture files (+ user.csv where noted).
TOKENIZER_CASES = [
    dict(name="test_tokenize_tokyo", line=11, input="東京都", user=False, ignore_space=False, max_grouping_len=0,
         tokens=[tok("東京都", (0, 3), (0, 9), F_TOKYOTO, -79 + 5320)]),
    dict(name="test_tokenize_kyotokyo", line=45, input="京都東京都京都", user=False, ignore_space=False,
         max_grouping_len=0,
         tokens=[tok("京都", (0, 2), (0, 6), F_KYOTO, -79 + 5293),
                 tok("東京都", (2, 5), (6, 15), F_TOKYOTO, -79 + 5293 + 569 + 5320),
                 tok("京都", (5, 7), (15, 21), F_KYOTO, -79 + 5293 + 569 + 5320 - 352 + 5293)]),
    dict(name="test_tokenize_kyotokyo_with_user", line=107, input="京都東京都京都", user=True, ignore_space=False,
         max_grouping_len=0,
         tokens=[tok("京都東京都", (0, 5), (0, 15), "カスタム名詞", -79 - 1000),
                 tok("京都", (5, 7), (15, 21), F_KYOTO, -79 - 1000 - 352 + 5293)]),
    dict(name="test_tokenize_tokyoto_with_space", line=154, input="東京 都", user=False, ignore_space=False,
         max_grouping_len=0,
         tokens=[tok("東京", (0, 2), (0, 6), F_TOKYO, -79 + 2816),
                 tok(" ", (2, 3), (6, 7), " ,空白,*,*,*,*,*, , ,*,A,*,*,*,*", -79 + 2816 - 390 - 20000),
                 tok("都", (3, 4), (7, 10), F_TO, -79 + 2816 - 390 - 20000 + 1134 + 2914)]),
    dict(name="test_tokenize_tokyoto_with_space_ignored", line=210, input="東京 都", user=False, ignore_space=True,
         max_grouping_len=0,
         tokens=[tok("東京", (0, 2), (0, 6), F_TOKYO, -79 + 2816),
                 tok("都", (3, 4), (7, 10), F_TO, -79 + 2816 - 390 + 2914)]),
    dict(name="test_tokenize_tokyoto_with_spaces_ignored", line=255, input="東京   都", user=False,
         ignore_space=True, max_grouping_len=0,
         tokens=[tok("東京", (0, 2), (0, 6), F_TOKYO, -79 + 2816),
                 tok("都", (5, 6), (9, 12), F_TO, -79 + 2816 - 390 + 2914)]),
    dict(name="test_tokenize_tokyoto_startswith_spaces_ignored", line=300, input="   東京都", user=False,
         ignore_space=True, max_grouping_len=0,
         tokens=[tok("東京都", (3, 6), (3, 12), F_TOKYOTO, -79 + 5320)]),
    dict(name="test_tokenize_tokyoto_endswith_spaces_ignored", line=334, input="東京都   ", user=False,
         ignore_space=True, max_grouping_len=0,
         tokens=[tok("東京都", (0, 3), (0, 9), F_TOKYOTO, -79 + 5320)]),
    dict(name="test_tokenize_kampersanda", line=368, input="kampersanda", user=False, ignore_space=False,
         max_grouping_len=0,
         tokens=[tok("kampersanda", (0, 11), (0, 11), F_ALPHA, 887 + 11633)]),
    dict(name="test_tokenize_kampersanda_with_user", line=399, input="kampersanda", user=True, ignore_space=False,
         max_grouping_len=0,
         tokens=[tok("kampersanda", (0, 11), (0, 11), "カスタム名詞", 887 - 2000)]),
    dict(name="test_tokenize_kampersanda_with_max_grouping", line=432, input="kampersanda", user=False,
         ignore_space=True, max_grouping_len=9,
         tokens=[tok("k", (0, 1), (0, 1), F_ALPHA, 887 + 11633),
                 tok("ampersanda", (1, 11), (1, 11), F_ALPHA, 887 + 11633 + 2341 + 11633)]),
    dict(name="test_tokenize_tokyoken", line=477, input="東京県に行く", user=False, ignore_space=False,
         max_grouping_len=0, num_tokens=4),
    dict(name="test_tokenize_kanjinumeric", line=495, input="一橋大学大学院", user=False, ignore_space=False,
         max_grouping_len=0,
         tokens=[tok("一橋大学大学院", (0, 7), (0, 21), "名詞,数,*,*,*,*,*")]),
    dict(name="test_tokenize_empty", line=520, input="", user=False, ignore_space=False, max_grouping_len=0,
         num_tokens=0),
]
# tests/tokenizer.rs:537 test_tokenize_repeat — one worker reused across sentences.
REPEAT_CASE = dict(name="test_tokenize_repeat", line=537,
                   sequence=[["東京に行く", 3], ["一橋大学大学院", 1], ["", 0], ["kampersanda", 1]])

# vibrato/src/tokenizer.rs:208-361 and token.rs:139-177 — inline mini-dictionaries.
MINI_LEX = "自然,0,0,1,sizen\n言語,0,0,4,gengo\n処理,0,0,3,shori\n自然言語,0,0,6,sizengengo\n言語処理,0,0,5,gengoshori"
MINI_CASES = [
    dict(name="tokenizer.rs::test_tokenize_1", line=209, lex=MINI_LEX, matrix="1 1\n0 0 0", char="DEFAULT 0 1 0",
         unk="DEFAULT,0,0,100,*", input="自然言語処理",
         tokens=[tok("自然", (0, 2), (0, 6), "sizen", 1), tok("言語処理", (2, 6), (6, 18), "gengoshori", 6)]),
    dict(name="tokenizer.rs::test_tokenize_2", line=252, lex=MINI_LEX, matrix="1 1\n0 0 0", char="DEFAULT 0 1 0",
         unk="DEFAULT,0,0,100,*", input="自然日本語処理",
         tokens=[tok("自然", (0, 2), (0, 6), "sizen", 1), tok("日本語処理", (2, 7), (6, 21), "*", 101)]),
    dict(name="tokenizer.rs::test_tokenize_3", line=295, lex=MINI_LEX, matrix="1 1\n0 0 0", char="DEFAULT 0 0 3",
         unk="DEFAULT,0,0,100,*", input="不自然言語処理",
         tokens=[tok("不自然", (0, 3), (0, 9), "*", 100), tok("言語処理", (3, 7), (9, 21), "gengoshori", 105)]),
    dict(name="tokenizer.rs::test_tokenize_empty", line=338, lex=MINI_LEX, matrix="1 1\n0 0 0",
         char="DEFAULT 0 0 3", unk="DEFAULT,0,0,100,*", input="", tokens=[]),
]

# tests/lexicon.rs:7-78, lexicon.rs:233-272 — common-prefix order and features.
LEXICON_CASES = dict(
    common_prefix_1=dict(line=8, input="東京都に行く",
                         matches=[[4, [7, 7, 4675], 1], [5, [6, 6, 2816], 2], [6, [6, 8, 5320], 3]]),
    common_prefix_2=dict(line=43, input="X", matches=[[w, [8, 8, -20000], 1] for w in range(40, 46)]),
    features=dict(line=60, items=[[0, "た,助動詞,*,*,*,助動詞-タ,終止形-一般,タ,た,*,A,*,*,*,*"],
                                  [2, "に,助詞,格助詞,*,*,*,*,ニ,に,*,A,*,*,*,*"],
                                  [39, " ,空白,*,*,*,*,*, , ,*,A,*,*,*,*"],
                                  [45, "X,名詞,固有名詞,地名,一般,*,*,X,X,*,A,*,*,*,*"]]),
    # lexicon.rs:233-272: WordMap::new(["東京","東京都","東京","京都"]) searched with "東京都"
    duplicate_surface=dict(line=234, words=["東京", "東京都", "東京", "京都"], input="東京都",
                           matches=[[0, 2], [2, 2], [1, 3]]),
    # lexicon.rs:274-329
    csv=dict(line=275,
             ok=[dict(data="自然,0,2,1,sizen\n言語,1,0,-4,gengo,げんご",
                      params=[[0, 2, 1], [1, 0, -4]], features=["sizen", "gengo,げんご"])],
             empty_surface=dict(data="自然,0,2,1,sizen\n,1,0,-4,gengo,げんご", n=1),
             errors=["自然,0,2", "自然,-2,2,1,a", "自然,2,-2,1,a", "自然,2,1,コスト,a"]),
)
# tests/connector.rs:5-14
MATRIX_CASES = dict(line=6, num_left=10, num_right=10, cost=[[0, 0, 0], [0, 1, 863], [1, 0, -3689], [9, 9, -2490]])

# SURVEY.md §8(d) worked counter values (derived by the survey's independent restatement; not pinned
# by the reference): input, user, ignore_space, max_grouping_len, walks,U,C,M,T,P,W,E,N,K,B_alg
COUNTER_CASES = [
    ["京都東京都京都", False, False, 0, 5, 21, 7, 14, 21, 14, 7, 10, 9, 3, 643],
    ["京都東京都京都", True, False, 0, 10, 21, 7, 27, 35, 16, 8, 12, 10, 2, 821],
    ["東京県に行く", False, False, 0, 5, 18, 6, 10, 14, 9, 8, 13, 10, 4, 600],
    ["kampersanda", False, True, 9, 2, 11, 11, 2, 2, 0, 2, 3, 4, 2, 225],
    ["東京   都", False, True, 0, 3, 12, 6, 6, 9, 6, 4, 6, 6, 2, 360],
]


def main():
    res = {}
    for fn in ["lex.csv", "matrix.def", "char.def", "unk.def", "user.csv"]:
        with open(os.path.join(REF, fn), encoding="utf-8", newline="") as f:
            res[fn] = f.read()
    out = dict(
        provenance="daac-tools/vibrato @ /root/reference (crate v0.5.2): vibrato/src/tests/resources/* and the "
                   "assertions of vibrato/src/tests/{tokenizer,lexicon,connector}.rs, tokenizer.rs:208-361, "
                   "lexicon.rs:233-329; generated by tests/golden/make_golden.py",
        resources=res,
        tokenizer_cases=TOKENIZER_CASES,
        repeat_case=REPEAT_CASE,
        mini_cases=MINI_CASES,
        lexicon_cases=LEXICON_CASES,
        matrix_cases=MATRIX_CASES,
        counter_cases=COUNTER_CASES,
    )
    with open(OUT, "w", encoding="utf-8") as f:
        json.dump(out, f, ensure_ascii=False, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
