"""An INDEPENDENT writer of vibrato's `.dic` stream, used to check the product's `Dictionary::read` against a second
derivation of the format — TEST INFRASTRUCTURE.

Everything here is derived from the reference's struct definitions as laid out in SURVEY.md Appendix A (container,
bincode rules, field order) and Appendix B (crawdad trie blob), and from the parsing rules of Appendix C — NOT from
vibrato_b200/csrc/host_dict.cpp, with which it shares no code, no helper and no data structure (plain dict / list /
struct.pack; a naive double-array builder).  Only MatrixConnector dictionaries are covered.

    dictionary.rs:27,43-51  magic + DictionaryInner      lexicon.rs:24-29      Lexicon
    map.rs:33-70            WordMap / builder             posting.rs:7-21       Postings
    trie.rs:14-27           trie blob                     param.rs:6-26         WordParam
    matrix_connector.rs     data/num_right/num_left       mapper.rs:9-12        ConnIdMapper
    character.rs:10-24,105  CharInfo bits / CharProperty  unknown.rs:21-27,63   UnkEntry / UnkHandler
    common.rs:5-9           bincode: little endian, fixed-width ints, u64 lengths, u8 Option tags, u32 enum tags
"""
import struct

MAGIC = b"VibratoTokenizer 0.5\n"
INVALID = 0x7FFFFFFF


def u64(v):
    return struct.pack("<Q", v)


def vec(items, fmt):
    return u64(len(items)) + b"".join(struct.pack("<" + fmt, x) for x in items)


def string(b):
    return u64(len(b)) + b


# ---- source files (Appendix C) ---------------------------------------------------------------------------------

def parse_lex(csv_text):
    """-> [(surface str, left, right, cost, feature bytes)]; rows with an empty surface are skipped (lexicon.rs:179-183)."""
    rows = []
    for line in csv_text.encode("utf-8").split(b"\n"):
        if line.endswith(b"\r"):
            line = line[:-1]
        if not line:
            continue
        assert not line.startswith(b'"'), "this writer does not handle quoted surfaces"
        parts = line.split(b",", 4)
        assert len(parts) == 5, line
        surface = parts[0].decode("utf-8")
        if surface == "":
            continue
        rows.append((surface, int(parts[1]), int(parts[2]), int(parts[3]), parts[4]))
    return rows


def parse_matrix(text):
    lines = [ln for ln in text.split("\n") if ln != ""]
    num_right, num_left = (int(x) for x in lines[0].split(" "))
    data = [0] * (num_right * num_left)
    for ln in lines[1:]:
        r, l, c = (int(x) for x in ln.split(" "))
        data[l * num_right + r] = c  # matrix_connector.rs:47
    return data, num_right, num_left


def parse_char_def(text):
    """-> (chr2inf list of 0x10000 u32, category names in id order)."""
    cats = {"DEFAULT": None}
    order = ["DEFAULT"]
    ranges = []
    for raw in text.split("\n"):
        ln = raw.strip()
        if not ln or ln.startswith("#"):
            continue
        if ln.startswith("0x"):
            ln = ln.split("#")[0].strip()
            cols = ln.split()
            if ".." in cols[0]:
                a, b = cols[0].split("..")
            else:
                a = b = cols[0]
            ranges.append((int(a, 16), int(b, 16), cols[1:]))
        else:
            cols = ln.split()
            name, invoke, group, length = cols[0], int(cols[1]), int(cols[2]), int(cols[3])
            if name not in cats:
                order.append(name)
            cats[name] = (invoke, group, length)
    ids = {n: i for i, n in enumerate(order)}

    def info(names):
        base = names[0]
        invoke, group, length = cats[base]
        idset = 0
        for n in names:
            idset |= 1 << ids[n]
        return idset | (ids[base] << 18) | (invoke << 26) | (group << 27) | (length << 28)  # character.rs:10-24

    table = [info(["DEFAULT"])] * 0x10000
    for a, b, names in ranges:  # later lines overwrite earlier ones
        v = info(names)
        for cp in range(a, b + 1):
            table[cp] = v
    return table, order


def parse_unk(text, cat_order):
    ids = {n: i for i, n in enumerate(cat_order)}
    per_cat = [[] for _ in cat_order]
    for line in text.encode("utf-8").split(b"\n"):
        if not line:
            continue
        parts = line.split(b",", 4)
        cid = ids[parts[0].decode()]
        per_cat[cid].append((cid, int(parts[1]), int(parts[2]), int(parts[3]), parts[4]))
    offsets, entries = [0], []
    for lst in per_cat:  # regrouped by category id, stable inside a category (unknown.rs:255-262)
        entries.extend(lst)
        offsets.append(len(entries))
    return offsets, entries


# ---- crawdad trie blob (Appendix B) --------------------------------------------------------------------------------

def build_trie_blob(keys_values):
    """keys_values: [(str key, u31 value)], keys unique.  A deliberately simple double array: children of a node sit
    at base ^ code, a key's end is the child with code 0 (flagged has_leaf on the parent), a leaf stores the value."""
    chars = sorted({c for k, _ in keys_values for c in k})
    code_of = {c: i + 1 for i, c in enumerate(chars)}  # code 0 is the end marker
    table_len = (max(ord(c) for c in chars) + 1) if chars else 0
    table = [0xFFFFFFFF] * table_len
    for c, code in code_of.items():
        table[ord(c)] = code
    alphabet = len(chars) + 1
    # logical trie
    root = {}
    for k, v in keys_values:
        node = root
        for c in k:
            node = node.setdefault(code_of[c], {})
        node[0] = v
    base, check, used = [INVALID], [INVALID], {0}

    def ensure(n):
        while len(base) <= n:
            base.append(INVALID)
            check.append(INVALID)

    pending = [(0, root)]
    while pending:
        idx, node = pending.pop()
        codes = sorted(node)
        b = 1
        while any((b ^ c) in used or (b ^ c) == 0 for c in codes):
            b += 1
        for c in codes:
            used.add(b ^ c)
            ensure(b ^ c)
        base[idx] = b  # not a leaf: MSB clear
        has_leaf = 0 in node
        for c in codes:
            child = b ^ c
            if c == 0:
                base[child] = 0x80000000 | node[0]  # leaf: MSB set, value below
                check[child] = idx
            else:
                check[child] = idx
                pending.append((child, node[c]))
        if has_leaf:
            check[idx] = (check[idx] & 0x7FFFFFFF) | 0x80000000 if idx else 0x80000000 | INVALID
    out = struct.pack("<I", table_len) + b"".join(struct.pack("<I", x) for x in table)
    out += struct.pack("<II", alphabet, len(base))
    out += b"".join(struct.pack("<II", base[i], check[i]) for i in range(len(base)))
    return out


# ---- sections -----------------------------------------------------------------------------------------------------

def lexicon_bytes(rows, lex_type):
    by_surface = {}
    for wid, (surface, _, _, _, _) in enumerate(rows):
        by_surface.setdefault(surface, []).append(wid)
    postings, kv = [], []
    for surface in sorted(by_surface, key=lambda s: [ord(c) for c in s]):  # BTreeMap<Vec<char>, _> order
        kv.append((surface, len(postings)))
        ids = by_surface[surface]
        postings.append(len(ids))
        postings.extend(ids)
    out = vec(list(build_trie_blob(kv)), "B")
    out += vec(postings, "I")
    out += u64(len(rows)) + b"".join(struct.pack("<HHh", l, r, c) for _, l, r, c, _ in rows)
    out += u64(len(rows)) + b"".join(string(f) for _, _, _, _, f in rows)
    out += struct.pack("<I", lex_type)
    return out


def dictionary_bytes(lex_csv, matrix_def, char_def, unk_def, user_csv=None, mapper=None):
    rows = parse_lex(lex_csv)
    data, num_right, num_left = parse_matrix(matrix_def)
    table, cat_order = parse_char_def(char_def)
    offsets, entries = parse_unk(unk_def, cat_order)
    out = bytearray(MAGIC)
    out += lexicon_bytes(rows, 0)
    if user_csv is None:
        out += b"\x00"
    else:
        out += b"\x01" + lexicon_bytes(parse_lex(user_csv), 1)
    out += struct.pack("<I", 0)  # ConnectorWrapper::Matrix
    out += vec(data, "h") + u64(num_right) + u64(num_left)
    if mapper is None:
        out += b"\x00"
    else:
        out += b"\x01" + vec(mapper[0], "H") + vec(mapper[1], "H")
    out += vec(table, "I")
    out += u64(len(cat_order)) + b"".join(string(c.encode()) for c in cat_order)
    out += vec(offsets, "Q")
    out += u64(len(entries)) + b"".join(struct.pack("<HHHh", c, l, r, w) + string(f) for c, l, r, w, f in entries)
    return bytes(out)


def strip_trie_blobs(stream, has_user):
    """The stream with the trie blobs cut out (two builders may lay a double array out differently): what is left
    must be byte-identical between writers."""
    assert stream.startswith(MAGIC)
    pos = len(MAGIC)
    pieces = [stream[:pos]]

    def skip_lexicon(p):
        (n,) = struct.unpack_from("<Q", stream, p)
        start = p + 8 + n  # after the trie blob
        q = start
        (m,) = struct.unpack_from("<Q", stream, q)
        q += 8 + 4 * m  # postings
        (m,) = struct.unpack_from("<Q", stream, q)
        q += 8 + 6 * m  # params
        (m,) = struct.unpack_from("<Q", stream, q)
        q += 8
        for _ in range(m):
            (ln,) = struct.unpack_from("<Q", stream, q)
            q += 8 + ln
        q += 4  # lex_type
        pieces.append(stream[start:q])
        return q

    pos = skip_lexicon(pos)
    tag = stream[pos]
    pieces.append(stream[pos:pos + 1])
    pos += 1
    if has_user:
        assert tag == 1
        pos = skip_lexicon(pos)
    pieces.append(stream[pos:])
    return b"".join(pieces)
