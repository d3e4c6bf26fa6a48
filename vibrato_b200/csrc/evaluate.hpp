// Library half of the reference's `evaluate` tool (evaluate/src/main.rs:61-138).
#pragma once

#include <cstdint>
#include <string_view>
#include <vector>

#include "engine.hpp"
#include "host_dict.hpp"

namespace vbt {

struct EvalCounts {
    uint64_t num_ref, num_sys, num_cor;  // evaluate/src/main.rs:80-82
};

// corpus: `surface\tfeature` lines with `EOS` between sentences (Corpus::from_reader, trainer/corpus.rs:78-121);
// feature_indices: the features compared, all when empty (evaluate/src/main.rs:33-37).  The engine is used as
// configured by the caller (the tool sets max_grouping_len and leaves ignore_space off, main.rs:72).
EvalCounts evaluate(const Dictionary& d, Engine& e, std::string_view corpus, const std::vector<uint64_t>& feature_indices);

}  // namespace vbt
