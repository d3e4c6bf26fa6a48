// `tokenize` look-alike: same flags and byte-identical output as the reference CLI
// (/root/reference/tokenize/src/main.rs:31-53 flags, :83-127 formats), tokenising on the GPU.
// Lines are shipped in batches; output order and content equal the per-line loop of main.rs:78-128.
#include <unistd.h>

#include <cstring>

#include "cli_common.hpp"

using namespace vibrato_b200;

static void usage() {
    std::fprintf(stderr,
                 "tokenize -i <system.dic.zst | mecab-source-dir> [-u user.csv] [-O mecab|wakati|detail] [-S] [-M n]\n"
                 "         [--format-on device|host]   (where the output text is built; default device)\n");
}

int main(int argc, char** argv) {
    std::string sysdic, userlex, mode = "mecab", format_on = "device";
    bool ignore_space = false;
    size_t max_grouping_len = 0;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() -> std::string {
            if (i + 1 >= argc) {
                usage();
                std::exit(2);
            }
            return argv[++i];
        };
        if (a == "-i" || a == "--sysdic") sysdic = next();
        else if (a == "-u" || a == "--userlex-csv") userlex = next();
        else if (a == "-O" || a == "--output-mode") mode = next();
        else if (a == "--format-on") format_on = next();
        else if (a == "-S" || a == "--ignore-space") ignore_space = true;
        else if (a == "-M" || a == "--max-grouping-len") max_grouping_len = std::stoull(next());
        else {
            usage();
            return 2;
        }
    }
    if (sysdic.empty() || (mode != "mecab" && mode != "wakati" && mode != "detail") ||
        (format_on != "device" && format_on != "host")) {
        if (!sysdic.empty()) std::fprintf(stderr, "Could not parse a mode\n");
        usage();
        return 2;
    }
    try {
        std::fprintf(stderr, "Loading the dictionary...\n");
        Dictionary dict = cli::load_dictionary(sysdic);
        if (!userlex.empty()) {
            std::string csv = cli::slurp(userlex);
            dict = std::move(dict).reset_user_lexicon_from_reader(&csv);
        }
        Tokenizer tokenizer = Tokenizer(std::move(dict)).ignore_space(ignore_space).max_grouping_len(max_grouping_len);
        if (format_on == "device") tokenizer.output_mode(mode);  // k_format_len / k_format_write build the text
        std::fprintf(stderr, "Ready to tokenize\n");
        const bool tty_out = isatty(STDOUT_FILENO), tty_in = isatty(STDIN_FILENO);
        const size_t batch_lines = tty_in ? 1 : 65536;
        std::string out;
        cli::Packed pk;
        std::string line;
        auto flush_batch = [&]() {
            if (pk.size() == 0) return;
            BatchResult r = tokenizer.tokenize_batch(pk.utf8.data(), pk.off.data(), pk.size());
            if (format_on == "device") {
                std::string_view text = r.text();
                std::fwrite(text.data(), 1, text.size(), stdout);
                if (tty_out) std::fflush(stdout);
                pk.clear();
                return;
            }
            const uint64_t* to = r.tok_offsets();
            for (uint64_t s = 0; s < pk.size(); ++s) {
                std::string_view sent(pk.utf8.data() + pk.off[s], pk.off[s + 1] - pk.off[s]);
                for (uint64_t k = to[s]; k < to[s + 1]; ++k) {
                    Token t(&tokenizer, sent, r.tokens()[k]);
                    if (mode == "mecab") {
                        out.append(t.surface()).append("\t").append(t.feature()).append("\n");
                    } else if (mode == "wakati") {
                        if (k != to[s]) out.append(" ");
                        out.append(t.surface());
                    } else {
                        out.append(t.surface()).append("\t").append(t.feature());
                        out.append("\tlex_type=").append(lex_type_name(t.lex_type()));
                        out.append("\tleft_id=").append(std::to_string(t.left_id()));
                        out.append("\tright_id=").append(std::to_string(t.right_id()));
                        out.append("\tword_cost=").append(std::to_string(t.word_cost()));
                        out.append("\ttotal_cost=").append(std::to_string(t.total_cost())).append("\n");
                    }
                }
                out.append(mode == "wakati" ? "\n" : "EOS\n");
                if (tty_out || out.size() > (1u << 20)) {
                    std::fwrite(out.data(), 1, out.size(), stdout);
                    if (tty_out) std::fflush(stdout);
                    out.clear();
                }
            }
            pk.clear();
        };
        while (cli::read_line(std::cin, line)) {
            pk.add(line);
            if (pk.size() >= batch_lines) flush_batch();
        }
        flush_batch();
        std::fwrite(out.data(), 1, out.size(), stdout);
        std::fflush(stdout);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "Error: %s\n", e.what());
        return 1;
    }
    return 0;
}
