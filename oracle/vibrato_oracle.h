/*
 * vibrato_oracle.h — CPU restatement of daac-tools/vibrato's tokenisation path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / `--impl reference` legs may load this library.  The
 * product path (vibrato_b200/) never links, imports or executes anything under oracle/.
 *
 * Parity status: the reference (Rust) cannot be built in this environment (no rustc /
 * cargo, dependency crates absent), so the oracle is pinned against the golden vectors
 * of the reference's own unit tests (vibrato/src/tests/tokenizer.rs, tests/lexicon.rs,
 * tests/connector.rs, tokenizer.rs:208-361, lexicon.rs:233-329) — see
 * tests/test_oracle_golden.py.  Third-party pieces restated from their published
 * behaviour: crawdad 0.3 common-prefix search semantics (vibrato call sites
 * dictionary/lexicon/map/trie.rs:49-56) and csv-core 0.1.10 field splitting
 * (dictionary/lexicon.rs:111-200).
 *
 * All file:line citations are relative to /root/reference/vibrato/src/.
 */
#ifndef VIBRATO_ORACLE_H
#define VIBRATO_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vo_dict vo_dict;
typedef struct vo_worker vo_worker;

/* Same 24-byte record as the product's vbt_token (include/vibrato_b200.h). */
typedef struct vo_token {
    uint32_t start_char; /* token.rs:21-24 range_char().start (start_word) */
    uint32_t end_char;   /* range_char().end */
    uint32_t start_byte; /* token.rs:28-32 range_byte() */
    uint32_t end_byte;
    uint32_t word_idx;   /* lex_type << 30 | word_id  (word_idx.rs:5-11, dictionary.rs:30-40) */
    int32_t total_cost;  /* token.rs:89-92: node.min_cost */
} vo_token;

/* Algorithmic-byte counters of SURVEY.md §8(d): U C M T P W E N K, then walks. */
enum { VO_CNT_U = 0, VO_CNT_C, VO_CNT_M, VO_CNT_T, VO_CNT_P, VO_CNT_W, VO_CNT_E, VO_CNT_N, VO_CNT_K,
       VO_CNT_WALKS, VO_NUM_COUNTERS };

/* SystemDictionaryBuilder::from_readers (dictionary/builder.rs:64-89). Returns NULL on error. */
vo_dict *vo_dict_from_mecab(const char *lex_csv, size_t lex_len, const char *matrix_def, size_t matrix_len,
                            const char *char_def, size_t char_len, const char *unk_def, size_t unk_len,
                            char *err, size_t errcap);
/* Same, but the connection matrix is given as a dense array data[left*num_right+right]
 * (matrix_connector.rs:11-15,79-85) — for synthetic dictionaries too large for a text matrix.def. */
vo_dict *vo_dict_from_parts(const char *lex_csv, size_t lex_len, const int16_t *matrix, uint32_t num_right,
                            uint32_t num_left, const char *char_def, size_t char_len, const char *unk_def,
                            size_t unk_len, char *err, size_t errcap);
/* Dictionary::reset_user_lexicon_from_reader (dictionary.rs:209-229); csv==NULL clears. 0 = ok. */
int vo_dict_set_user_csv(vo_dict *d, const char *csv, size_t len, char *err, size_t errcap);
void vo_dict_free(vo_dict *d);

/* Dictionary::word_feature (dictionary.rs:108-114). */
const char *vo_dict_feature(const vo_dict *d, uint32_t word_idx, size_t *len);
/* Dictionary::word_param (dictionary.rs:98-104). */
int vo_dict_word_param(const vo_dict *d, uint32_t word_idx, uint16_t *left, uint16_t *right, int16_t *cost);
/* MatrixConnector::cost (matrix_connector.rs:121-124). */
int32_t vo_dict_conn_cost(const vo_dict *d, uint16_t right_id, uint16_t left_id);
uint32_t vo_dict_num_left(const vo_dict *d);
uint32_t vo_dict_num_right(const vo_dict *d);
uint32_t vo_dict_num_words(const vo_dict *d, int lex_type);
/* CharProperty::char_info (character.rs:112-116): packed CharInfo(u32). */
uint32_t vo_dict_char_info(const vo_dict *d, uint32_t cp);
/* Lexicon::common_prefix_iterator (lexicon.rs:33-46). Writes up to cap (word_id,end_char) pairs. */
size_t vo_dict_common_prefix(const vo_dict *d, int lex_type, const uint32_t *chars, size_t n, uint32_t *word_ids,
                             uint32_t *end_chars, size_t cap);

/* Tokenizer::new(dict).ignore_space(..)?.max_grouping_len(..).new_worker()  (tokenizer.rs:26-84).
 * Returns NULL (err set) when ignore_space is requested and SPACE is undefined (tokenizer.rs:44-49). */
vo_worker *vo_worker_new(const vo_dict *d, int ignore_space, uint64_t max_grouping_len, char *err, size_t errcap);
void vo_worker_free(vo_worker *w);
/* Worker::reset_sentence + tokenize (worker.rs:34-55). Input must be valid UTF-8. Returns num_tokens. */
size_t vo_worker_tokenize(vo_worker *w, const char *utf8, size_t len);
/* Tokens in sentence order (worker.rs:65-68 already reversed). Valid until the next tokenize call. */
const vo_token *vo_worker_tokens(const vo_worker *w);
/* Same as vo_worker_tokenize but also accumulates the §8(d) counters into cnt[VO_NUM_COUNTERS]. */
size_t vo_worker_tokenize_counted(vo_worker *w, const char *utf8, size_t len, uint64_t *cnt);

/* Batch driver: sentences i = utf8[off[i]..off[i+1]); n_threads workers over a static partition.
 * tok_off (n+1 entries) and *toks (malloc'ed, caller frees with vo_free) receive all tokens when
 * toks != NULL; otherwise only counts are produced (the benchmark protocol,
 * benchmark/src/main.rs:53-65: reset_sentence + tokenize + num_tokens).
 * cnt (optional) receives summed counters.  Returns total tokens. */
uint64_t vo_tokenize_batch(const vo_dict *d, int ignore_space, uint64_t max_grouping_len, const char *utf8,
                           const uint64_t *off, uint64_t n, int n_threads, uint64_t *tok_off, vo_token **toks,
                           uint64_t *cnt);
/* Timed variant for the CPU baseline: runs `runs` passes, returns seconds of the whole call. */
double vo_benchmark(const vo_dict *d, int ignore_space, uint64_t max_grouping_len, const char *utf8,
                    const uint64_t *off, uint64_t n, int n_threads, int runs, uint64_t *n_words);
void vo_free(void *p);

/* Worker::init_connid_counter / update_connid_counts (worker.rs:77-94 -> Lattice::add_connid_counts,
 * lattice.rs:170-183) summed over a batch: lid_count[num_left], rid_count[num_right] (ConnIdCounter,
 * mapper.rs:87-104).  Returns 0 on success. */
int vo_connid_counts_batch(const vo_dict *d, int ignore_space, uint64_t max_grouping_len, const char *utf8,
                           const uint64_t *off, uint64_t n, int n_threads, uint64_t *lid_count,
                           uint64_t *rid_count);
/* Dictionary::map_connection_ids_from_iter (dictionary.rs:245-259; ConnIdMapper::from_iter/parse
 * mapper.rs:40-80; MatrixConnector::map_connection_ids matrix_connector.rs:99-116). 0 = ok. */
int vo_dict_map_connection_ids(vo_dict *d, const uint16_t *lmap, size_t n_lmap, const uint16_t *rmap, size_t n_rmap,
                               char *err, size_t errcap);

/* SystemDictionaryBuilder::from_readers_with_bigram_info (dictionary/builder.rs:111-148) with
 * dual_connector = false: the connection costs come from a RawConnector
 * (connector/raw_connector.rs:22-160 over raw_connector/scorer.rs:103-267) built from
 * bigram.right / bigram.left / bigram.cost. */
vo_dict *vo_dict_from_bigram(const char *lex_csv, size_t lex_len, const char *bigram_right, size_t right_len,
                             const char *bigram_left, size_t left_len, const char *bigram_cost, size_t cost_len,
                             const char *char_def, size_t char_len, const char *unk_def, size_t unk_len, char *err,
                             size_t errcap);
/* Same with dual_connector = true: DualConnector::from_readers (connector/dual_connector.rs:155-213) — a reduced
 * matrix over all but eight feature templates plus an 8-lane raw term (cost :269-280, map_connection_ids
 * :227-266).  Ties of the greedy template choice (:27-70) follow a HashSet's iteration order upstream; here
 * the highest template index among the ties is dropped (see vibrato_oracle.c). */
vo_dict *vo_dict_from_bigram_dual(const char *lex_csv, size_t lex_len, const char *bigram_right, size_t right_len,
                                  const char *bigram_left, size_t left_len, const char *bigram_cost, size_t cost_len,
                                  const char *char_def, size_t char_len, const char *unk_def, size_t unk_len, char *err,
                                  size_t errcap);
/* Scorer built from (key1, key2, cost) triples with ScorerBuilder::insert/build (scorer.rs:110-168), then
 * Scorer::accumulate_cost over two feature-id rows (scorer.rs:255-267) — for the reference's scorer vectors. */
int32_t vo_scorer_accumulate(const uint32_t *triples, size_t n_triples, const uint32_t *keys1, const uint32_t *keys2,
                             size_t n_keys);

/* std::str::from_utf8 validity (the check `stdin.lines()` applies before the hot path). 1 = valid. */
int vo_utf8_valid(const char *s, size_t len);

#ifdef __cplusplus
}
#endif
#endif
