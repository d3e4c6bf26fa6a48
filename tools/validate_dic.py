#!/usr/bin/env python3
"""Validation kit for a RELEASED vibrato dictionary (`system.dic` or `system.dic.zst`) — SURVEY.md §8 rows a13 / f-1.

No released dictionary exists in the build environment, so the `.dic` byte layout and the
crawdad trie blob (Appendix B) are pinned only by tests/dic_format.py, a second derivation from the reference's
struct definitions.  The day a real dictionary is mounted, this script is the check:

    python tools/validate_dic.py /path/to/ipadic-mecab-2_7_0/system.dic.zst [--user user.csv] [--no-gpu]

  1. load     Dictionary::read through the C ABI (vbt_dict_from_zstd_file / vbt_dict_from_bytes); the loader itself
              cross-checks trie values against the postings and the word count (host_dict.cpp read_lexicon)
  2. audit    every key of the trie is enumerated and looked up again; every word id is named by exactly one key
              (vbt_dict_audit)
  3. rewrite  Dictionary::write(Dictionary::read(x)) == x, byte for byte (after zstd decompression)
  4. README   the known answers of /root/reference/README.md:66-90 and :124-140 (ipadic-mecab-2.7.0): MeCab-format
              output of `本とカレーの街神保町へようこそ。`, its wakati form, and `mens second bag` with and without
              `-S -M 24` — through the GPU tokenizer and the device-side output stage (needs a GPU; --no-gpu skips)

Prints a pass/fail table and exits non-zero on any failure.  Step 4 only applies to ipadic-mecab-2.7.0; for other
dictionaries it prints the tokenisation for inspection.
"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vibrato_b200 as vb  # noqa: E402

README_MECAB = """本\t名詞,一般,*,*,*,*,本,ホン,ホン
と\t助詞,並立助詞,*,*,*,*,と,ト,ト
カレー\t名詞,固有名詞,地域,一般,*,*,カレー,カレー,カレー
の\t助詞,連体化,*,*,*,*,の,ノ,ノ
街\t名詞,一般,*,*,*,*,街,マチ,マチ
神保\t名詞,固有名詞,地域,一般,*,*,神保,ジンボウ,ジンボー
町\t名詞,接尾,地域,*,*,*,町,マチ,マチ
へ\t助詞,格助詞,一般,*,*,*,へ,ヘ,エ
ようこそ\t感動詞,*,*,*,*,*,ようこそ,ヨウコソ,ヨーコソ
。\t記号,句点,*,*,*,*,。,。,。
EOS
"""  # README.md:72-83
README_WAKATI = "本 と カレー の 街 神保 町 へ ようこそ 。\n"  # README.md:89
README_SPACES = """mens\t名詞,固有名詞,組織,*,*,*,*
 \t記号,空白,*,*,*,*,*
second\t名詞,固有名詞,組織,*,*,*,*
 \t記号,空白,*,*,*,*,*
bag\t名詞,固有名詞,組織,*,*,*,*
EOS
"""  # README.md:124-130
README_MECAB_COMPAT = """mens\t名詞,固有名詞,組織,*,*,*,*
second\t名詞,一般,*,*,*,*,*
bag\t名詞,固有名詞,組織,*,*,*,*
EOS
"""  # README.md:136-140 (-S -M 24)


def zstd_decompress(data):
    """One-shot decompression through libzstd.so.1 (the image has neither the `zstd` tool nor a Python binding)."""
    z = ctypes.CDLL("libzstd.so.1")
    z.ZSTD_getFrameContentSize.restype = ctypes.c_ulonglong
    z.ZSTD_getFrameContentSize.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
    z.ZSTD_decompress.restype = ctypes.c_size_t
    z.ZSTD_decompress.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
    z.ZSTD_isError.argtypes = [ctypes.c_size_t]
    n = z.ZSTD_getFrameContentSize(data, len(data))
    if n in (2**64 - 1, 2**64 - 2):  # error / unknown size: fall back to a generous bound
        n = len(data) * 40
    buf = ctypes.create_string_buffer(n)
    got = z.ZSTD_decompress(buf, n, data, len(data))
    if z.ZSTD_isError(got):
        raise RuntimeError("zstd: cannot decompress the file in one shot")
    return buf.raw[:got]


def zstd_compress(data, level=3):
    z = ctypes.CDLL("libzstd.so.1")
    z.ZSTD_compressBound.restype = ctypes.c_size_t
    z.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
    z.ZSTD_compress.restype = ctypes.c_size_t
    z.ZSTD_compress.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int]
    cap = z.ZSTD_compressBound(len(data))
    buf = ctypes.create_string_buffer(cap)
    got = z.ZSTD_compress(buf, cap, data, len(data), level)
    return buf.raw[:got]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("dic")
    ap.add_argument("--user", help="user.csv to attach (Dictionary::reset_user_lexicon_from_reader)")
    ap.add_argument("--no-gpu", action="store_true")
    ap.add_argument("--ipadic", action="store_true", help="force the README comparison (default: when the path says ipadic)")
    a = ap.parse_args()
    rows = []

    def row(name, ok, detail=""):
        rows.append((name, ok, detail))
        print(f"[{'PASS' if ok else 'FAIL'}] {name}{': ' + detail if detail else ''}", flush=True)

    # 1. load
    try:
        if a.dic.endswith(".zst"):
            d = vb.Dictionary.from_zstd_file(a.dic)
            raw = zstd_decompress(open(a.dic, "rb").read())
        else:
            raw = open(a.dic, "rb").read()
            d = vb.Dictionary.read(raw)
        sh = d.shape()
        row("load", True, f"{sh['n_system']} words, {sh['num_left']} x {sh['num_right']} connection ids, "
                          f"{sh['n_unknown']} unknown entries")
    except Exception as e:  # noqa: BLE001
        row("load", False, str(e))
        return finish(rows)
    if a.user:
        d = d.reset_user_lexicon_from_reader(open(a.user, "rb").read())
    # 2. audit
    for lex, name in ((0, "system"), (1, "user")):
        if lex == 1 and not a.user:
            continue
        au = d.audit(lex)
        ok = au["words"] == au["listed"] and au["keys_not_found"] == 0 and au["words_unlisted_or_twice"] == 0
        row(f"audit {name} lexicon", ok, f"{au['keys']} keys, {au['words']} words, longest key {au['longest_key']} chars, "
                                         f"{au['keys_not_found']} keys lost, {au['words_unlisted_or_twice']} words off")
    # 3. rewrite
    if not a.user:
        out = d.write()
        row("write(read(x)) == x", out == raw, f"{len(raw)} bytes" if out == raw else
            f"{len(out)} vs {len(raw)} bytes, first difference at {next((i for i, (p, q) in enumerate(zip(out, raw)) if p != q), min(len(out), len(raw)))}")
    # 4. README
    if a.no_gpu:
        row("README known answers", True, "skipped (--no-gpu)")
        return finish(rows)
    is_ipadic = a.ipadic or "ipadic" in os.path.basename(os.path.dirname(os.path.abspath(a.dic))).lower() or \
        "ipadic" in os.path.basename(a.dic).lower()

    def text_of(tok, sentence, mode):
        tok.output_mode(mode)
        _, text = tok.tokenize_batch([sentence]).text()
        return text.decode("utf-8")

    cases = [("本とカレーの街神保町へようこそ。", "mecab", False, 0, README_MECAB),
             ("本とカレーの街神保町へようこそ。", "wakati", False, 0, README_WAKATI),
             ("mens second bag", "mecab", False, 0, README_SPACES),
             ("mens second bag", "mecab", True, 24, README_MECAB_COMPAT)]
    for sentence, mode, ignore_space, max_group, want in cases:
        try:
            tok = vb.Tokenizer.new(d)
            if ignore_space:
                tok = tok.ignore_space(True).max_grouping_len(max_group)
            got = text_of(tok, sentence, mode)
        except Exception as e:  # noqa: BLE001
            row(f"tokenize `{sentence}` ({mode})", False, str(e))
            continue
        if is_ipadic:
            row(f"README `{sentence}` -O {mode}{' -S -M 24' if ignore_space else ''}", got == want,
                "" if got == want else "got:\n" + got)
        else:
            print(f"--- `{sentence}` -O {mode}{' -S -M 24' if ignore_space else ''}\n{got}", flush=True)
    return finish(rows)


def finish(rows):
    bad = [r for r in rows if not r[1]]
    print(f"\n{len(rows) - len(bad)} / {len(rows)} checks passed")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
