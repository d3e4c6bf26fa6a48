import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vibrato_b200 as vb
from vibrato_b200 import synth
sd = synth.make_dictionary("synth-small")
d = vb.SystemDictionaryBuilder.from_readers(sd.lex_csv, sd.matrix, sd.char_def, sd.unk_def)
utf8, off = synth.make_corpus(sd, 6001, seed=3, log_uniform=(1, 300), unk_frac=0.1, space_frac=0.02)
for lanes, sort, chunk, smem, dual, cnt in [(8,0,1000,0,0,0),(8,0,1000,0,0,1),(8,0,1000,0,1,0),(8,0,1000,1,0,0),(8,0,1000,1,1,1)]:
    tok = vb.Tokenizer.new(d)
    tok.set_option("lanes_per_sentence", lanes); tok.set_option("sort_by_length", sort)
    tok.set_option("chunk_sentences", chunk); tok.set_option("smem_rows", smem); tok.set_option("dual_stream", dual)
    tok.set_counting(bool(cnt))
    print("config", lanes, sort, chunk, smem, dual, cnt, flush=True)
    res = tok.tokenize_batch(utf8=utf8, byte_offsets=off)
    print("  ok", res.n_tokens, flush=True)
