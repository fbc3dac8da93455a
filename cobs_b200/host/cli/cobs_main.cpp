// cobs -- command line front-end of the B200 query path.  `cobs query` is a drop-in for the
// reference's subtool (src/cobs.cpp:410-527): same flags and defaults, same stdout format
// ("doc\tscore" per result, "*comment\t<count>" per FASTA record), TIMER line on stderr.
// FASTA records are searched in GPU batches instead of one at a time.
#include <cobs/query/classic_index/mmap_search_file.hpp>
#include <cobs/query/classic_search.hpp>
#include <cobs/query/compact_index/mmap_search_file.hpp>
#include <cobs/settings.hpp>
#include <cobs/util/error_handling.hpp>
#include <cobs/util/file.hpp>

#include <cobsgpu.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <future>
#include <iostream>
#include <iterator>
#include <map>
#include <memory>
#include <random>
#include <string>
#include <vector>

namespace {

void usage_query(std::ostream& os) {
    os << "Usage: cobs query [options] [query]\n\n"
          "Parameters:\n"
          "  [query]              the text sequence to search for\n"
          "Options:\n"
          "  -i, --index <path>   path to index file(s), may be given several times\n"
          "  -f, --file <path>    query (fasta) file to process\n"
          "  -t, --threshold <x>  threshold in percentage of terms in query matching, default: 0.8\n"
          "  -l, --limit <n>      number of results to return, default: all\n"
          "      --load-complete  accepted for compatibility (the index always lives in HBM)\n"
          "  -T, --threads <n>    accepted for compatibility (the search runs on the GPU)\n"
          "      --gpus <n>       shard every index over n GPUs along the document axis\n"
          "      --device <d>     first CUDA device to use, default: 0\n"
          "      --batch <n>      FASTA records per GPU batch, default: 16384\n";
}

struct Batch {
    std::vector<std::string> comments, queries;
};

// COBS_CLI_TRACE=1: wall seconds per stage of `cobs query -f`, one extra line on stderr
struct CliTrace {
    bool on = std::getenv("COBS_CLI_TRACE") != nullptr;
    double open = 0, parse = 0, wait = 0, search = 0, format = 0, write = 0;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    static double since(std::chrono::steady_clock::time_point t) {
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t).count();
    }
    void print() const {
        if (!on) return;
        std::fprintf(stderr, "CLI open=%.3f parse=%.3f wait_for_worker=%.3f | worker: search=%.3f "
                     "format=%.3f write=%.3f | wall=%.3f\n", open, parse, wait, search, format, write, since(t0));
    }
} g_trace;

// "<comment>\t<count>\n" + "<doc>\t<score>\n" per result, formatted into one buffer per batch
// and written with a single write (operator<< per field costs more than the GPU search, and so
// does snprintf: the digits are produced by hand)
void format_batch(const Batch& b, const std::vector<std::vector<cobs::SearchResult> >& results,
                  std::string& out) {
    out.clear();
    char num[24];
    auto put_num = [&](unsigned long v) {
        char* end = num + sizeof(num);
        char* p = end;
        do {
            *--p = char('0' + v % 10);
            v /= 10;
        } while (v != 0);
        out.append(p, size_t(end - p));
    };
    for (size_t i = 0; i < results.size(); ++i) {
        out += b.comments[i];
        out += '\t';
        put_num(results[i].size());
        out += '\n';
        for (const auto& res : results[i]) {
            out += res.doc_name;
            out += '\t';
            put_num(res.score);
            out += '\n';
        }
    }
}

// one batch on its way through the stages: parsed records, then their results
struct Stage {
    Batch b;
    std::vector<std::vector<cobs::SearchResult> > results;
};

void write_stage(const Stage& st) {
    auto t = std::chrono::steady_clock::now();
    std::string out;
    format_batch(st.b, st.results, out);
    g_trace.format += CliTrace::since(t);
    t = std::chrono::steady_clock::now();
    std::cout.write(out.data(), std::streamsize(out.size()));
    g_trace.write += CliTrace::since(t);
}

// same record splitting as process_query (src/cobs.cpp:425-462).  The FASTA file is processed as
// a three-stage pipeline: this thread parses batch i+2 while one worker searches batch i+1 on the
// GPU and another formats and writes the lines of batch i -- in order, so stdout is what the
// reference prints.
void process_query(cobs::Search& s, double threshold, unsigned num_results,
                   const std::string& query_line, const std::string& query_file, size_t batch) {
    if (!query_line.empty()) {
        std::vector<cobs::SearchResult> result;
        s.search(query_line, result, threshold, num_results);
        for (const auto& res : result) std::cout << res.doc_name << '\t' << res.score << '\n';
    }
    else if (!query_file.empty()) {
        std::ifstream qf(query_file);
        std::string line, query, comment;
        Batch parsing;
        std::future<void> searcher, writer;   // `writer` is only touched by the searcher tasks
        auto t_parse = std::chrono::steady_clock::now();
        auto dispatch = [&] {
            g_trace.parse += CliTrace::since(t_parse);
            auto t_wait = std::chrono::steady_clock::now();
            if (searcher.valid()) searcher.get();   // one search at a time, in order
            g_trace.wait += CliTrace::since(t_wait);
            if (!parsing.queries.empty()) {
                auto st = std::make_shared<Stage>();
                st->b.comments.swap(parsing.comments);
                st->b.queries.swap(parsing.queries);
                searcher = std::async(std::launch::async, [&s, &writer, st, threshold, num_results] {
                    auto t = std::chrono::steady_clock::now();
                    s.search_batch(st->b.queries, st->results, threshold, num_results);
                    g_trace.search += CliTrace::since(t);
                    if (writer.valid()) writer.get();   // batch i-1 is written before batch i
                    writer = std::async(std::launch::async, [st] { write_stage(*st); });
                });
            }
            t_parse = std::chrono::steady_clock::now();
        };
        auto push = [&] {
            parsing.comments.push_back(comment);
            parsing.queries.push_back(query);
            if (parsing.queries.size() >= batch) dispatch();
        };
        while (std::getline(qf, line)) {
            if (line.empty()) continue;
            if (line[0] == '>' || line[0] == ';') {
                if (!query.empty()) push();
                line[0] = '*';
                query.clear();
                comment = line;
            }
            else {
                query += line;
            }
        }
        if (!query.empty()) push();
        dispatch();
        auto t_wait = std::chrono::steady_clock::now();
        if (searcher.valid()) searcher.get();
        if (writer.valid()) writer.get();
        g_trace.wait += CliTrace::since(t_wait);
    }
    else {
        cobs::die_with_message("Pass a verbatim query or a query file.");
    }
    std::cout.flush();
    s.timer().print("search");
    g_trace.print();
}

int query(int argc, char** argv) {
    std::vector<std::string> index_files;
    std::string query, query_file;
    double threshold = 0.8;
    unsigned num_results = 0;
    size_t batch = 16384;

    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto value = [&](const char* name) -> std::string {
            if (i + 1 >= argc) {
                std::cerr << "Error: option " << name << " requires an argument!\n\n";
                usage_query(std::cerr);
                std::exit(-1);
            }
            return argv[++i];
        };
        if (a == "-i" || a == "--index") index_files.push_back(value("-i"));
        else if (a == "-f" || a == "--file") query_file = value("-f");
        else if (a == "-t" || a == "--threshold") threshold = std::atof(value("-t").c_str());
        else if (a == "-l" || a == "--limit") num_results = unsigned(std::strtoul(value("-l").c_str(), nullptr, 10));
        else if (a == "--load-complete") cobs::gopt_load_complete_index = true;
        else if (a == "-T" || a == "--threads") cobs::gopt_threads = std::strtoul(value("-T").c_str(), nullptr, 10);
        else if (a == "--gpus") cobs::gopt_gpus = unsigned(std::max(1, std::atoi(value("--gpus").c_str())));
        else if (a == "--device") cobs::gopt_gpu_device = std::atoi(value("--device").c_str());
        else if (a == "--batch") batch = std::max<size_t>(1, std::strtoul(value("--batch").c_str(), nullptr, 10));
        else if (a == "-h" || a == "--help") {
            usage_query(std::cout);
            return -1;
        }
        else if (!a.empty() && a[0] == '-' && a.size() > 1) {
            std::cerr << "Error: unknown option \"" << a << "\".\n\n";
            usage_query(std::cerr);
            return -1;
        }
        else if (query.empty()) query = a;
        else {
            std::cerr << "Error: unexpected extra argument \"" << a << "\".\n\n";
            usage_query(std::cerr);
            return -1;
        }
    }

    std::vector<std::shared_ptr<cobs::IndexSearchFile> > indices;
    for (auto& path : index_files) {
        if (cobs::file_has_header<cobs::ClassicIndexHeader>(path))
            indices.push_back(std::make_shared<cobs::ClassicIndexMMapSearchFile>(path));
        else if (cobs::file_has_header<cobs::CompactIndexHeader>(path))
            indices.push_back(std::make_shared<cobs::CompactIndexMMapSearchFile>(path));
        else
            cobs::die_with_message("Could not open index path \"" + path + "\"");
    }
    cobs::ClassicSearch s(indices);
    g_trace.open = CliTrace::since(g_trace.t0);
    process_query(s, threshold, num_results, query, query_file, batch);
    return 0;
}

// `cobs benchmark-fpr`: the reference's micro-benchmark of the query path
// (src/cobs.cpp:605-730): mt19937-seeded random queries of num_kmers + 30 bp, warm-up, then a
// loop of search() with threshold 0 / all results, and a "RESULT name=benchmark ..." line.
// --batch N (extension) sends N queries per GPU batch instead of one search() per query.
int benchmark_fpr(int argc, char** argv) {
    std::string in_file;
    unsigned num_kmers = 1000, num_queries = 10000, num_warmup = 100;
    bool fpr_dist = false;
    size_t seed = std::random_device { } ();
    size_t batch = 1;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto value = [&](const char* name) -> std::string {
            if (i + 1 >= argc) {
                std::cerr << "Error: option " << name << " requires an argument!\n";
                std::exit(-1);
            }
            return argv[++i];
        };
        if (a == "-k" || a == "--num-kmers") num_kmers = unsigned(std::strtoul(value("-k").c_str(), nullptr, 10));
        else if (a == "-q" || a == "--queries") num_queries = unsigned(std::strtoul(value("-q").c_str(), nullptr, 10));
        else if (a == "-w" || a == "--warmup") num_warmup = unsigned(std::strtoul(value("-w").c_str(), nullptr, 10));
        else if (a == "-d" || a == "--dist") fpr_dist = true;
        else if (a == "--seed") seed = std::strtoull(value("--seed").c_str(), nullptr, 10);
        else if (a == "--batch") batch = std::max<size_t>(1, std::strtoul(value("--batch").c_str(), nullptr, 10));
        else if (a == "--gpus") cobs::gopt_gpus = unsigned(std::max(1, std::atoi(value("--gpus").c_str())));
        else if (a == "--device") cobs::gopt_gpu_device = std::atoi(value("--device").c_str());
        else if (!a.empty() && a[0] == '-') {
            std::cerr << "Error: unknown option \"" << a << "\".\n";
            return -1;
        }
        else if (in_file.empty()) in_file = a;
    }
    if (in_file.empty()) {
        std::cerr << "Usage: cobs benchmark-fpr [-k num_kmers] [-q queries] [-w warmup] [-d] "
                     "[--seed s] [--batch n] <in_file>\n";
        return -1;
    }
    // same generator and draw order as the reference: one mt19937 stream, warm-up queries
    // first (cobs::random_sequence_rng, cobs/util/misc.hpp:31-38)
    std::mt19937 rng(seed);
    auto random_sequence = [&](size_t size) {
        static const char basepairs[4] = { 'A', 'C', 'G', 'T' };
        std::string r;
        for (size_t j = 0; j < size; ++j) r += basepairs[rng() % 4];
        return r;
    };
    std::vector<std::string> warmup_queries, queries;
    for (unsigned i = 0; i < num_warmup; ++i) warmup_queries.push_back(random_sequence(num_kmers + 30));
    for (unsigned i = 0; i < num_queries; ++i) queries.push_back(random_sequence(num_kmers + 30));

    cobs::ClassicSearch s(std::make_shared<cobs::ClassicIndexMMapSearchFile>(in_file));
    std::vector<std::vector<cobs::SearchResult> > results;
    auto run = [&](const std::vector<std::string>& qs, std::map<uint32_t, uint64_t>* counts) {
        for (size_t b = 0; b < qs.size(); b += batch) {
            std::vector<std::string> part(qs.begin() + b, qs.begin() + std::min(qs.size(), b + batch));
            s.search_batch(part, results);
            if (counts)
                for (auto& res : results)
                    for (auto& r : res) (*counts)[r.score]++;
        }
    };
    run(warmup_queries, nullptr);
    s.timer().reset();
    std::map<uint32_t, uint64_t> counts;
    auto t0 = std::chrono::steady_clock::now();
    run(queries, fpr_dist ? &counts : nullptr);
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    cobs::Timer t = s.timer();
    std::cout << "RESULT"
              << " name=benchmark "
              << " index=" << in_file
              << " kmer_queries=" << (queries.empty() ? 0 : queries[0].size() - 30)
              << " queries=" << queries.size()
              << " warmup=" << warmup_queries.size()
              << " results=" << (results.empty() ? 0 : results.back().size())
              << " sse2=off"
              << " aio=off"
              << " t_hashes=" << t.get("hashes")
              << " t_io=" << t.get("io")
              << " t_and=" << t.get("and rows")
              << " t_add=" << t.get("add rows")
              << " t_sort=" << t.get("sort results")
              << " gpu=on batch=" << batch << " t_wall=" << wall
              << std::endl;
    for (const auto& c : counts)
        std::cout << "RESULT name=benchmark_fpr fpr=" << c.first << " dist=" << c.second << std::endl;
    return 0;
}

// `cobs classic-construct <input> <out_file>`: the reference's subtool (src/cobs.cpp:117-190) for
// FASTA and plain-text documents, built on the device (cobsgpu_construct_classic) and written in
// the reference's file format -- byte-identical to what the reference writes from the same files.
// One document per file, named after the file without its extension, ordered by path.
int classic_construct(int argc, char** argv) {
    std::string input, out_file;
    unsigned num_hashes = 1, term_size = 31;
    double fpr = 0.3;
    bool no_canonicalize = false, clobber = false;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto value = [&](const char* name) -> std::string {
            if (i + 1 >= argc) {
                std::cerr << "Error: option " << name << " requires an argument!\n";
                std::exit(-1);
            }
            return argv[++i];
        };
        if (a == "-h" || a == "--num-hashes") num_hashes = unsigned(std::strtoul(value("-h").c_str(), nullptr, 10));
        else if (a == "-f" || a == "--false-positive-rate") fpr = std::atof(value("-f").c_str());
        else if (a == "-k" || a == "--term-size") term_size = unsigned(std::strtoul(value("-k").c_str(), nullptr, 10));
        else if (a == "--no-canonicalize") no_canonicalize = true;
        else if (a == "-C" || a == "--clobber") clobber = true;
        else if (a == "--device") cobs::gopt_gpu_device = std::atoi(value("--device").c_str());
        else if (a == "-T" || a == "--threads" || a == "-m" || a == "--memory" || a == "--tmp-path" ||
                 a == "--file-type") value(a.c_str());      // accepted for compatibility
        else if (a == "--continue" || a == "--keep-temporary") { }
        else if (!a.empty() && a[0] == '-') {
            std::cerr << "Error: unknown option \"" << a << "\".\n";
            return -1;
        }
        else if (input.empty()) input = a;
        else if (out_file.empty()) out_file = a;
    }
    if (input.empty() || out_file.empty()) {
        std::cerr << "Usage: cobs classic-construct [-h num_hashes] [-f fpr] [-k term_size] "
                     "[--no-canonicalize] [-C] <input dir or file> <out_file>\n";
        return -1;
    }
    namespace fs = cobs::fs;
    if (fs::exists(out_file) && !clobber)
        cobs::die_with_message("Output file exists, will not overwrite without --clobber.");
    auto is_fasta = [](const std::string& ext) {
        for (const char* e : { ".fa", ".fasta", ".fna", ".ffn", ".faa", ".frn" })
            if (ext == e) return true;
        return false;
    };
    std::vector<fs::path> files;
    auto consider = [&](const fs::path& p) {
        const std::string ext = p.extension().string();
        if (is_fasta(ext) || ext == ".txt") files.push_back(p);
    };
    if (fs::is_directory(input)) {
        for (auto& e : fs::recursive_directory_iterator(input))
            if (e.is_regular_file()) consider(e.path());
    }
    else consider(input);
    std::sort(files.begin(), files.end());
    if (files.empty()) cobs::die_with_message("No FASTA or text documents found in \"" + input + "\"");

    // documents -> sequences (records of a FASTA file the way the reference walks it,
    // cobs/fasta_file.hpp:156-182: '>' / ';' lines and empty lines end a record)
    std::vector<std::string> names;
    std::string sequences;
    std::vector<uint64_t> seq_offsets { 0 };
    std::vector<uint32_t> seq_doc;
    for (size_t d = 0; d < files.size(); ++d) {
        names.push_back(files[d].stem().string());
        std::ifstream in(files[d], std::ios::binary);
        std::string data((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
        auto end_record = [&] {
            if (sequences.size() > seq_offsets.back()) {
                seq_offsets.push_back(sequences.size());
                seq_doc.push_back(uint32_t(d));
            }
        };
        if (!is_fasta(files[d].extension().string())) {
            sequences += data;
            end_record();
            continue;
        }
        size_t pos = 0;
        while (pos <= data.size()) {
            size_t nl = data.find('\n', pos);
            if (nl == std::string::npos) nl = data.size();
            if (nl == pos || data[pos] == '>' || data[pos] == ';') end_record();
            else sequences.append(data, pos, nl - pos);
            pos = nl + 1;
        }
        end_record();
    }
    std::vector<const char*> name_ptrs;
    for (auto& n : names) name_ptrs.push_back(n.c_str());
    cobsgpu_construct_desc desc;
    std::memset(&desc, 0, sizeof(desc));
    desc.struct_size = sizeof(desc);
    desc.term_size = term_size;
    desc.canonicalize = no_canonicalize ? 0 : 1;
    desc.num_hashes = num_hashes;
    desc.signature_size = 0;
    desc.false_positive_rate = fpr;
    desc.n_docs = uint32_t(names.size());
    desc.n_seqs = uint32_t(seq_doc.size());
    desc.doc_names = name_ptrs.data();
    desc.sequences = sequences.data();
    desc.seq_offsets = seq_offsets.data();
    desc.seq_doc = seq_doc.data();
    desc.device = cobs::gopt_gpu_device;
    cobsgpu_index* ix = nullptr;
    if (cobsgpu_construct_classic(&desc, &ix) != COBSGPU_OK)
        cobs::die_with_message(std::string("construction failed: ") + cobsgpu_last_error());
    const int rc = cobsgpu_index_save(ix, out_file.c_str());
    cobsgpu_index_info info;
    cobsgpu_index_get_info(ix, &info);
    std::cerr << "classic index: " << info.n_docs << " documents, signature size "
              << cobsgpu_index_signature_size(ix, 0) << ", " << info.num_hashes << " hashes -> "
              << out_file << std::endl;
    cobsgpu_index_close(ix);
    if (rc != COBSGPU_OK) cobs::die_with_message(std::string("could not write index: ") + cobsgpu_last_error());
    return 0;
}

void usage(const char* prog) {
    std::cout << "(Co)mpact (B)it-Sliced (S)ignature Index for Genome Search -- B200 query path\n\n"
              << "Usage: " << prog << " <subtool> ...\n\n"
              << "Available subtools:\n"
              << "  query          query an index (classic or compact) on the GPU\n"
              << "  benchmark-fpr  the reference's query micro-benchmark (classic index)\n"
              << "  classic-construct  build a classic index from FASTA / text documents on the GPU\n"
              << "  version        print version\n\n"
              << "Compact construction and the other subtools of the reference are not part of\n"
              << "this build; indices written by the reference are read as they are.\n";
}

} // namespace

int main(int argc, char** argv) {
    if (argc < 2) {
        usage(argv[0]);
        return 0;
    }
    const std::string tool = argv[1];
    try {
        if (tool == "query") return query(argc - 1, argv + 1);
        if (tool == "benchmark-fpr") return benchmark_fpr(argc - 1, argv + 1);
        if (tool == "classic-construct") return classic_construct(argc - 1, argv + 1);
        if (tool == "version") {
            std::cout << "COBS B200 query path, C ABI version " << cobsgpu_version() << std::endl;
            return 0;
        }
    }
    catch (std::exception& e) {
        // reference: src/cobs.cpp:1070-1076
        std::cerr << "EXCEPTION: " << e.what() << std::endl;
        return -1;
    }
    std::cout << "Unknown subtool \"" << tool << "\"\n";
    usage(argv[0]);
    return 0;
}
