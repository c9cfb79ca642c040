// kasa_b200 host CLI — `build` and `identify` with kASA's flags, files and outputs (main.cpp:303-586 of the
// reference for the flag names; formats per SURVEY.md §3.4 / Compare.hpp:1485-1890, 3466-3665).  FASTA/FASTQ(.gz)
// parsing, content file, scoring and writers stay on the host; all k-mer work goes through libkasa_b200.so.
#include <dirent.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <chrono>
#include <iostream>
#include <memory>
#include <unordered_set>

#include "../../include/kasa_b200.h"
#include "kasa_host.hpp"

using namespace kasa_host;

namespace {

// Growable array in page-locked host memory (kasa_host_alloc): the staging buffers of a batch.  Only the few vector operations
// the batch loop uses; the contents are trivially copyable.
template <typename T>
struct Pinned {
    T *p = nullptr;
    size_t n = 0, cap = 0;
    Pinned() = default;
    Pinned(const Pinned &) = delete;
    Pinned &operator=(const Pinned &) = delete;
    Pinned(Pinned &&o) noexcept : p(o.p), n(o.n), cap(o.cap) { o.p = nullptr; o.n = o.cap = 0; }
    ~Pinned() { kasa_host_free(p); }
    void reserve(size_t want)
    {
        if (want <= cap) return;
        size_t c = std::max<size_t>(want, cap * 2 + 64);
        void *q = nullptr;
        if (kasa_host_alloc(&q, c * sizeof(T)) != KASA_OK) throw std::runtime_error("page-locked host memory could not be allocated");
        if (n) std::memcpy(q, p, n * sizeof(T));
        kasa_host_free(p);
        p = static_cast<T *>(q); cap = c;
    }
    void clear() { n = 0; }
    void push_back(const T &v) { reserve(n + 1); p[n++] = v; }
    void append(const T *src, size_t k) { reserve(n + k); std::memcpy(p + n, src, k * sizeof(T)); n += k; }
    void assign(size_t k, const T &v) { reserve(k); std::fill(p, p + k, v); n = k; }
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
    T *data() { return p; }
    const T *data() const { return p; }
};

// regular entries of a directory, sorted by name
std::vector<std::string> list_directory(const std::string &path)
{
    std::vector<std::string> out;
    DIR *d = opendir(path.c_str());
    if (!d) throw std::runtime_error("Input folder not found: " + path);
    while (dirent *e = readdir(d)) {
        const std::string n = e->d_name;
        if (n == "." || n == ".." || e->d_type == DT_DIR) continue;
        out.push_back(n);
    }
    closedir(d);
    std::sort(out.begin(), out.end());
    return out;
}

struct Args {
    std::string mode, content, index, input, in1, in2, profile, rtt, tmp, alphabet_file, alphabet_id, out, first_index, second_index, deleted, content_out;
    int k_high = 12, k_low = 7, highest_k = 12, beasts = 3, device = 0, gpus = 1;
    bool six = false, one = false, verbose = false, exact = false, show = false;
    float threshold = 0.f;
    std::string format = "json";
    uint32_t batch_reads = 1u << 20;
    uint64_t batch_bases = 0;
};

[[noreturn]] void usage()
{
    std::cout << "kasa_b200 <build|identify> -c <content> -d <index> [-i <input> | -1 <fq> -2 <fq>] [-p <profile.csv>] [-q <per-read out>]\n"
                 "kasa_b200 identify_multiple -c <content> -d <index> -i <folder> [-p <profile prefix>] [-q <per-read prefix>] (identify's flags)\n"
                 "kasa_b200 update -c <content> -d <index> -i <new sequences> [-o <out>] (the content file names the new accessions)\n"
                 "kasa_b200 --parameters <parameters.yaml>     (kASA's parameter file instead of flags)\n"
                 "kasa_b200 half -c <content> -d <index> -o <out>     (6-byte records, the upper six letters in <out>_trie)\n"
                 "kasa_b200 redundancy -c <content> -d <index> [-v]\n"
                 "kasa_b200 trie -d <index> | getFrequency -c <content> -d <index> | delete -c <content> -d <index> -o <out> -l <delnodes.dmp>\n"
                 "          | merge -c <content> --firstIndex <a> --secondIndex <b> -o <out>     (indexes sharing one content file)\n"
                 "          [-k <high> <low>] [--kH <12|25>] [--six | --one] [--json|--jsonl|--tsv|--kraken] [-b <n>] [--threshold <f>] [-v]\n"
                 "          [-a <gc.prt> <id>]\n"
                 "          [--exact-scores] [--device <n>] [--gpus <n>] [--batch-reads <n>] [--batch-bases <n>]   (-m -n -t -r -x are accepted and ignored)\n";
    std::exit(0);
}

// `--parameters <file>` (main.cpp:264, Utilities::readParametersFromYaml Utilities.hpp:1114-): the flat "Key: value" lines of kASA's
// parameters.yaml; section headers and comments are skipped, quotes dropped.  Keys of features this front end does not have are
// accepted when they are off and refused when they are on.
Args parse_yaml(const std::string &path)
{
    std::ifstream f(path);
    if (!f) throw std::runtime_error("Config file not found!");
    Args a;
    std::string line;
    auto truth = [](const std::string &v) { return v == "true" || v == "True" || v == "1"; };
    bool one = false, three = false, six = false;
    while (std::getline(f, line)) {
        const size_t b = line.find_first_not_of(" \t");
        if (b == std::string::npos || line[b] == '#') continue;
        const size_t colon = line.find(':', b);
        if (colon == std::string::npos) continue;
        std::string key = line.substr(b, colon - b), val = line.substr(colon + 1);
        key.erase(std::remove(key.begin(), key.end(), '"'), key.end());
        val.erase(std::remove(val.begin(), val.end(), '"'), val.end());
        val.erase(0, std::min(val.find_first_not_of(" \t"), val.size()));
        while (!val.empty() && (val.back() == '\r' || val.back() == ' ' || val.back() == '\t')) val.pop_back();
        if (key == "Mode") a.mode = val;
        else if (key == "Index") a.index = val;
        else if (key == "ContentFile") a.content = val;
        else if (key == "kHigh") { if (!val.empty()) a.k_high = std::stoi(val); }
        else if (key == "kLow") { if (!val.empty()) a.k_low = std::stoi(val); }
        else if (key == "FilePathForTemporaryFiles") a.tmp = val;
        else if (key == "Verbose") a.verbose = truth(val);
        else if (key == "AlphabetFile") a.alphabet_file = val;
        else if (key == "AlphabetIndex") a.alphabet_id = val;
        else if (key == "InputFileOrFolder") a.input = val;
        else if (key == "PairedEnd-First") a.in1 = val;
        else if (key == "PairedEnd-Second") a.in2 = val;
        else if (key == "One") one = truth(val);
        else if (key == "Three") three = truth(val);
        else if (key == "Six") six = truth(val);
        else if (key == "ProfileOutputfile") a.profile = val;
        else if (key == "ReadIDtoTaxIDOutputfile") a.rtt = val;
        else if (key == "ReadIDtoTaxIDOutputFormat") { if (!val.empty()) a.format = val; }
        else if (key == "NumberOfTaxaPerRead") { if (!val.empty()) { a.beasts = std::stoi(val); if (a.beasts == 0) a.beasts = 1; } }
        else if (key == "ThresholdForScore") { if (!val.empty()) a.threshold = std::stof(val); }
        else if (key == "FileWithDeletedTaxa") a.deleted = val;
        else if (key == "ContentFile-First" || key == "ContentFile-Second") { if (a.content.empty()) a.content = val; }
        else if (key == "ContentFile-Out") a.content_out = val;
        else if (key == "FirstOldIndex") a.first_index = val;
        else if (key == "SecondOldIndex") a.second_index = val;
        else if (key == "NewIndex") a.out = val;
        else if (key == "AlreadyTranslated" || key == "UniqueKmersOnly" || key == "PrintCoverage" || key == "Coherence" || key == "Spaced" || key == "Visualize" ||
                 key == "Debug" || key == "TaxIDsAreStrings") {
            if (truth(val)) throw std::runtime_error("parameter " + key + " is not supported by kasa_b200");
        }
        else if (key == "Filter") { if (!val.empty() && val != "_ _") throw std::runtime_error("parameter Filter is not supported by kasa_b200"); }
        // NumberOfThreads, AvailableRAMinGB, CallIndex, UseRAMOnly, TaxonomicLevel, AccessionToTaxIDFileOrFolder, TaxonomyFolder, IGotSpace,
        // ErrorThreshold, CoherenceThreshold, Gzip, ShrinkingStrategy, ShrinkPercentage, SpacedMaskIdx: nothing to set here
    }
    if (one + three + six > 1) throw std::runtime_error("Only one of the frames One / Three / Six can be true");
    a.one = one; a.six = six;
    if (a.alphabet_id.empty() || a.alphabet_file.empty()) { a.alphabet_file.clear(); a.alphabet_id.clear(); }
    if (a.k_high > 12) a.highest_k = 25;
    if (a.mode.empty()) throw std::runtime_error("the parameter file names no Mode");
    return a;
}

Args parse(int argc, char **argv)
{
    Args a;
    if (argc < 2) usage();
    if (std::string(argv[1]) == "--parameters") {
        if (argc < 3) throw std::runtime_error("missing value after --parameters");
        a = parse_yaml(argv[2]);
        if (argc > 3 && std::string(argv[3]) == "--show") a.show = true;          // prints what was read (tests)
        return a;
    }
    a.mode = argv[1];
    for (int i = 2; i < argc; ++i) {
        const std::string p = argv[i];
        auto next = [&]() -> std::string { if (i + 1 >= argc) throw std::runtime_error("missing value after " + p); return argv[++i]; };
        if (p == "-c" || p == "--content") a.content = next();
        else if (p == "-d" || p == "--database") a.index = next();
        else if (p == "-i" || p == "--input") a.input = next();
        else if (p == "-o" || p == "--outgoing") a.out = next();
        else if (p == "--firstIndex") a.first_index = next();
        else if (p == "--secondIndex") a.second_index = next();
        else if (p == "-l" || p == "--deleted") a.deleted = next();
        else if (p == "-co") a.content_out = next();
        else if (p == "-c1" || p == "-c2") { const std::string c = next(); if (a.content.empty()) a.content = c; }
        else if (p == "-1") a.in1 = next();
        else if (p == "-2") a.in2 = next();
        else if (p == "-p" || p == "--profile") a.profile = next();
        else if (p == "-q" || p == "--rtt") a.rtt = next();
        else if (p == "-t" || p == "--temp") a.tmp = next();
        else if (p == "-k") { a.k_high = std::stoi(next()); a.k_low = std::stoi(next()); }
        else if (p == "--kH") a.highest_k = std::stoi(next()) > 12 ? 25 : 12;
        else if (p == "--kL") (void)next();
        else if (p == "-a" || p == "--alphabet") { a.alphabet_file = next(); a.alphabet_id = next(); }
        else if (p == "-b" || p == "--beasts") { a.beasts = std::stoi(next()); if (a.beasts == 0) a.beasts = 1; }
        else if (p == "--threshold") a.threshold = std::stof(next());
        else if (p == "--six") a.six = true;
        else if (p == "--three") a.six = false;
        else if (p == "--one") a.one = true;
        else if (p == "--json") a.format = "json";
        else if (p == "--jsonl") a.format = "jsonl";
        else if (p == "--tsv") a.format = "tsv";
        else if (p == "--kraken") a.format = "kraken";
        else if (p == "-v" || p == "--verbose") a.verbose = true;
        else if (p == "--exact-scores") a.exact = true;
        else if (p == "--device") a.device = std::stoi(next());
        else if (p == "--gpus") a.gpus = std::max(1, std::stoi(next()));
        else if (p == "--batch-reads") a.batch_reads = (uint32_t)std::stoul(next());
        else if (p == "--batch-bases") a.batch_bases = std::stoull(next());
        else if (p == "-m" || p == "-n" || p == "-x") (void)next();
        else if (p == "-r" || p == "--ram") {}
        else if (p == "-h" || p == "--help") usage();
        else throw std::runtime_error("unknown parameter " + p);
    }
    return a;
}

void check(int rc, kasa_ctx *ctx)
{
    if (rc != KASA_OK) throw std::runtime_error(kasa_last_error(ctx));
}

// `-a <file> <id>`: the translation table `id` of an NCBI gc.prt file replaces the codons it names in the built-in table
// (kASA::setCodonTable, kASA.hpp:579-615: same lookup of "  id <n> ,", same five lines, '*' becomes '[').
// Returns false (built-in table stays, with the reference's warning) when the id is not in the file.
bool load_codon_table(const std::string &file, const std::string &id, uint8_t table[512])
{
    kasa_builtin_codon_table(table);
    std::ifstream f(file);
    std::string line;
    bool found = false;
    while (std::getline(f, line))
        if (line.find("  id " + id + " ,") != std::string::npos) { found = true; break; }
    if (!found) { std::cerr << "WARNING: codon table not found in file. Using built-in." << std::endl; return false; }
    std::string aa, dummy, b1, b2, b3;
    std::getline(f, aa); std::getline(f, dummy); std::getline(f, b1); std::getline(f, b2); std::getline(f, b3);
    size_t pa = aa.find_first_of('"') + 1, pb = b1.find_first_of("TGCA");
    for (; pb != std::string::npos && pb < b1.size() && pb < b2.size() && pb < b3.size() && pa < aa.size(); ++pb, ++pa)
        table[((b1[pb] & 14) << 5) | ((b2[pb] & 14) << 2) | ((b3[pb] & 14) >> 1)] = aa[pa] == '*' ? (uint8_t)'[' : (uint8_t)aa[pa];
    return true;
}

// ------------------------------------------------------------------------------------------------ index files
enum { kWriteIndex = 1, kWriteTrie = 2, kWriteFrequency = 4, kWriteAll = 7 };

// kasa_build_finish + the files of an index: <out>, <out>_info.txt (Build.hpp:466-470), <out>_trie, <out>_trie.txt (Trie.hpp:387-388),
// <out>_f.txt (kASA.hpp:549-570)
void finish_and_write(kasa_ctx *ctx, const std::string &out, int K, const Content &content, const std::string &content_path, int what)
{
    uint64_t n = 0, nt = 0;
    check(kasa_build_finish(ctx, &n, &nt), ctx);
    const size_t rec = K > 12 ? 20 : 12, block = K > 12 ? 2048000 : 2101248;
    if (what & (kWriteIndex | kWriteTrie)) {
        std::vector<char> entries(n * rec), trie(nt * 12);
        check(kasa_build_fetch(ctx, entries.data(), trie.data()), ctx);
        if (what & kWriteIndex) {
            write_padded(out, entries.data(), entries.size(), block);
            std::ofstream f(out + "_info.txt");
            f << n;
            if (K > 12) f << std::endl << 128;
        }
        if (what & kWriteTrie) {
            write_padded(out + "_trie", trie.data(), trie.size(), 2101248);
            std::ofstream f(out + "_trie.txt");
            f << nt;
        }
    }
    if (what & kWriteFrequency) {
        std::vector<uint64_t> freq(content.taxids.size() * (size_t)K);
        check(kasa_build_frequency(ctx, content.taxids.data(), (uint32_t)content.taxids.size(), freq.data()), ctx);
        std::ofstream f(out + "_f.txt");
        std::ifstream cf(content_path);
        f << "non_unique";
        for (int k = 0; k < K; ++k) f << "\t" << freq[k];
        f << std::endl;
        std::string line;
        for (size_t j = 1; j < content.taxids.size() && std::getline(cf, line);) {
            if (line.empty()) continue;
            f << split(line, '\t')[0] << "\t" << freq[j * K];
            for (int k = 1; k < K; ++k) f << "\t" << freq[j * K + k];
            f << std::endl;
            ++j;
        }
    }
}

// the records of an index file: <idx>_info.txt = count [\n 128 | 3] (main.cpp:982-991)
std::vector<char> read_index_records(const std::string &path, int *bits)
{
    std::ifstream info(path + "_info.txt");
    if (!info) throw std::runtime_error("Info file for this index can not be found!");
    uint64_t n = 0;
    std::string tok;
    info >> n;
    *bits = 64;
    if (info >> tok) {
        if (tok == "128") *bits = 128;
        else if (tok == "3") throw std::runtime_error("Halved indices cannot be modified in this way. Sorry...");      // main.cpp:787
    }
    std::vector<char> rec(n * (*bits == 128 ? 20 : 12));
    std::ifstream f(path, std::ios::binary);
    if (!f || !f.read(rec.data(), (std::streamsize)rec.size())) throw std::runtime_error("Library file does not exist or is shorter than its info file says: " + path);
    return rec;
}

// `half` (main.cpp:1465-1474, Shrink::putHalfInTrie Shrink.hpp:78-143): one streaming pass over the index file on the host — nothing to
// compute.  <out> = 6-byte records {u32 lower six letters, u16 position of the taxid in the content file}, without the records whose lower
// six letters are all '^'; <out>_trie counts the kept records per upper six letters (the last run one short unless it is 1, as the
// reference writes it); <out>_info.txt = "<n>\n3"; <out>_f.txt is the old frequency file (Shrink.hpp:441-452).
int run_half(const Args &a)
{
    if (a.out.empty() || a.out == a.index) throw std::runtime_error("Paths and names of input and output are the same!");      // main.cpp:1466
    int bits = 64;
    const std::vector<char> rec = read_index_records(a.index, &bits);
    if (bits != 64) throw std::runtime_error("Only 64-bit indices can be halved");
    Content content = load_content(a.content.empty() ? a.index + "_content.txt" : a.content);
    if (content.taxids.size() > 65535) throw std::runtime_error("Too many taxa for a halved index (2^16 - 1 at most)");
    std::unordered_map<uint32_t, uint16_t> pos;
    for (size_t i = 0; i < content.taxids.size(); ++i) pos[content.taxids[i]] = (uint16_t)i;
    const uint64_t n = rec.size() / 12, low_mask = (1ull << 30) - 1;
    std::vector<char> out6, trie;
    out6.reserve(n * 6);
    uint64_t count = 0, current = 0;
    bool first = true;
    auto put_run = [&](uint64_t prefix, uint64_t c) {
        const uint32_t p32 = (uint32_t)prefix;
        trie.insert(trie.end(), reinterpret_cast<const char *>(&c), reinterpret_cast<const char *>(&c) + 8);
        trie.insert(trie.end(), reinterpret_cast<const char *>(&p32), reinterpret_cast<const char *>(&p32) + 4);
    };
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t k;
        uint32_t t;
        std::memcpy(&k, rec.data() + i * 12, 8);
        std::memcpy(&t, rec.data() + i * 12 + 8, 4);
        if (first) { current = k >> 30; first = false; }
        if ((k & low_mask) == 1039104990ull) continue;                     // "^^^^^^"
        auto it = pos.find(t);
        if (it == pos.end()) throw std::runtime_error("a taxid of the index is missing from the content file: " + std::to_string(t));
        const uint32_t lo = (uint32_t)(k & low_mask);
        out6.insert(out6.end(), reinterpret_cast<const char *>(&lo), reinterpret_cast<const char *>(&lo) + 4);
        out6.insert(out6.end(), reinterpret_cast<const char *>(&it->second), reinterpret_cast<const char *>(&it->second) + 2);
        ++count;
        if ((k >> 30) != current) { put_run(current, count - 1); current = k >> 30; count = 1; }
    }
    put_run(current, count != 1 ? count - 1 : 1);
    write_padded(a.out, out6.data(), out6.size(), 2101248);
    { std::ofstream f(a.out + "_info.txt"); f << out6.size() / 6 << std::endl << 3; }
    write_padded(a.out + "_trie", trie.data(), trie.size(), 2101248);
    { std::ofstream f(a.out + "_trie.txt"); f << trie.size() / 12; }
    {
        std::ifstream src(a.index + "_f.txt", std::ios::binary);
        std::ofstream dst(a.out + "_f.txt", std::ios::binary);
        dst << src.rdbuf();
    }
    return 0;
}

// `redundancy` (main.cpp:1364-1420, Shrink::histogram Shrink.hpp:33-70): how many taxa the k-mers of an index have — one streaming pass
// on the host.  hist[c] = k-mers listed with c taxa; the cutoff is the smallest c up to which 99 % of the index's records lie.
int run_redundancy(const Args &a)
{
    int bits = 64;
    const std::vector<char> rec = read_index_records(a.index, &bits);         // (refuses halved indices)
    const size_t rb = bits == 128 ? 20 : 12, kb = rb - 4;
    const uint64_t n = rec.size() / rb;
    Content content = load_content(a.content.empty() ? a.index + "_content.txt" : a.content);
    const uint32_t n_tax = (uint32_t)content.taxids.size();                      // content lines + 1, as the reference counts them
    std::vector<uint64_t> hist(n_tax + 2, 0);
    uint64_t unique = 0;
    uint32_t counter = 0;
    for (uint64_t i = 0; i < n; ++i) {
        if (i == 0 || std::memcmp(rec.data() + i * rb, rec.data() + (i - 1) * rb, kb) == 0) ++counter;
        else { hist[std::min<uint32_t>(counter, n_tax + 1)]++; counter = 1; ++unique; }
    }
    hist[std::min<uint32_t>(counter, n_tax + 1)]++;
    if (a.verbose) std::cout << "Number of unique k-mers: " << unique << std::endl << "Histogram\nFrequency Counts Percentage" << std::endl;
    double percentage = 0.0;
    uint32_t cutoff = 0;
    for (uint32_t i = 1; i < n_tax + 1; ++i) {
        if (a.verbose && hist[i]) std::cout << i << " " << hist[i] << " " << 100.0 * (double)hist[i] * i / (double)n << std::endl;
        percentage += (double)hist[i] * i / (double)n;
        if (percentage >= 0.99 && cutoff == 0) cutoff = i;
    }
    if (cutoff == 1) std::cout << "OUT: 99% of the k-mers in your index have only one taxon. Using unique frequencies makes sense." << std::endl;
    else if (cutoff < 4) std::cout << "OUT: 99% of the k-mers in your index have " << cutoff << " or less taxa. Using unique frequencies could make sense." << std::endl;
    else std::cout << "OUT: 99% of the k-mers in your index have " << cutoff << " or less taxa. You should consider looking at the non-unique frequencies as well." << std::endl;
    return 0;
}

// trie / getFrequency / delete / merge (main.cpp:1422-1458, 1336-1362, 770-805, 877-977): an index file's records join a build on the
// device (kasa_build_push_index), whose finish is the merge, the prefix table and the frequencies.
//   trie          rewrites <index>_trie, <index>_trie.txt
//   getFrequency  rewrites <index>_f.txt
//   delete        <out> = <index> without the records of the taxids listed in the first column of -l (Update.hpp:28-94)
//   merge         <out> = --firstIndex + --secondIndex, both built with the content file -c (Read.hpp:3180-3243; the reference writes no
//                 <out>_info.txt and, for want of it, a frequency file of zeros: here both are written properly)
int run_index_mode(const Args &a)
{
    const bool merge = a.mode == "merge", del = a.mode == "delete";
    const std::string first = merge ? a.first_index : a.index;
    if (first.empty()) throw std::runtime_error("No index given!");
    if ((merge || del) && a.out.empty()) throw std::runtime_error("No output file given!");
    if (merge && a.first_index == a.second_index) throw std::runtime_error("-d and -i must point to different indices!");
    if (merge && (a.out == a.first_index || a.out == a.second_index)) throw std::runtime_error("You can't overwrite indices (yet)!");
    int bits = 64, bits2 = 64;
    std::vector<char> rec = read_index_records(first, &bits), rec2;
    if (merge) {
        rec2 = read_index_records(a.second_index, &bits2);
        if (bits != bits2) throw std::runtime_error("Indices are not of the same format!");
    }
    const int K = bits == 128 ? 25 : 12;
    const size_t rb = bits == 128 ? 20 : 12;
    Content content;
    const std::string content_path = !a.content.empty() ? a.content : first + "_content.txt";
    if (a.mode != "trie") content = load_content(content_path);
    if (del) {
        std::ifstream f(a.deleted);
        if (!f) throw std::runtime_error("delnodes.dmp not found");
        std::unordered_set<uint32_t> gone;
        std::string line;
        while (std::getline(f, line))
            if (!line.empty()) gone.insert((uint32_t)std::stoul(split(line, '\t')[0]));
        size_t w = 0;                                      // the taxid is the last four bytes of a record (packedPairs.hpp:100-166)
        for (size_t r = 0; r + rb <= rec.size(); r += rb) {
            uint32_t t;
            std::memcpy(&t, rec.data() + r + rb - 4, 4);
            if (gone.count(t)) continue;
            if (w != r) std::memmove(rec.data() + w, rec.data() + r, rb);
            w += rb;
        }
        rec.resize(w);
    }
    kasa_params p{};
    p.device = a.device; p.highest_k = K; p.k_high = K; p.k_low = std::min(7, K); p.frames = 3;
    p.score_mode = KASA_SCORE_NONE;
    const uint64_t total = (rec.size() + rec2.size()) / rb;
    p.max_batch_kmers = std::min<uint64_t>(std::max<uint64_t>(total + (1u << 20), 1u << 22), (1ull << 32) - (1u << 20));
    p.max_batch_seqs = 1u << 10;
    p.max_batch_bases = 1u << 20;
    kasa_ctx *ctx = nullptr;
    check(kasa_ctx_create(&p, &ctx), nullptr);
    check(kasa_build_push_index(ctx, rec.data(), rec.size() / rb, bits), ctx);
    if (merge) check(kasa_build_push_index(ctx, rec2.data(), rec2.size() / rb, bits), ctx);
    const std::string out = (merge || del) ? a.out : a.index;
    const int what = a.mode == "trie" ? kWriteTrie : a.mode == "getFrequency" ? kWriteFrequency : kWriteAll;
    finish_and_write(ctx, out, K, content, content_path, what);
    if (merge && !a.content_out.empty() && a.content_out != content_path) {        // -co: the merged content file = the shared one
        std::ifstream src(content_path, std::ios::binary);
        std::ofstream dst(a.content_out, std::ios::binary);
        dst << src.rdbuf();
    }
    kasa_ctx_destroy(ctx);
    return 0;
}

// ------------------------------------------------------------------------------------------------ build
// `update` (main.cpp:699-768 with the content file given, Update.hpp:99-: UpdateFromFasta) is a build whose brick starts with the
// records of the old index: <out> (default: the index itself) = <index> + the sequences of -i.
int run_build(const Args &a)
{
    const bool update = a.mode == "update";
    int K = a.highest_k, old_bits = 64;
    std::vector<char> old_records;
    if (update) {
        old_records = read_index_records(a.index, &old_bits);
        K = old_bits == 128 ? 25 : 12;
    }
    const std::string out_path = update && !a.out.empty() ? a.out : a.index;
    Content content = load_content(a.content.empty() && update ? a.index + "_content.txt" : a.content);
    const std::string fasta = slurp(a.input);
    kasa_params p{};
    if (a.one && a.six) throw std::runtime_error("--one and --six exclude each other");     // main.cpp:618
    p.device = a.device; p.highest_k = K; p.k_high = K; p.k_low = std::min(7, K); p.frames = a.one ? 1 : (a.six ? 6 : 3);
    p.score_mode = KASA_SCORE_NONE;
    uint8_t codon[512];
    if (!a.alphabet_file.empty() && load_codon_table(a.alphabet_file, a.alphabet_id, codon)) p.codon_table = codon;
    const uint64_t windows = (uint64_t)fasta.size() * (a.six ? 2 : 1) + (1u << 20) + old_records.size() / (old_bits == 128 ? 20 : 12);
    p.max_batch_kmers = std::min<uint64_t>(std::max<uint64_t>(windows * 2, 1u << 22), (1ull << 32) - (1u << 20));
    p.max_batch_seqs = 1u << 20;
    p.max_batch_bases = 256ull << 20;
    kasa_ctx *ctx = nullptr;
    check(kasa_ctx_create(&p, &ctx), nullptr);
    if (update) {
        check(kasa_build_push_index(ctx, old_records.data(), old_records.size() / (old_bits == 128 ? 20 : 12), old_bits), ctx);
        std::vector<char>().swap(old_records);
    }
    std::string chars;
    std::vector<uint64_t> starts;
    std::vector<uint32_t> lens, tax;
    auto flush = [&]() {
        if (starts.empty()) return;
        kasa_seq_batch sb{(uint32_t)starts.size(), KASA_MEM_HOST, chars.data(), starts.data(), lens.data(), tax.data()};
        check(kasa_build_push(ctx, &sb), ctx);
        chars.clear(); starts.clear(); lens.clear(); tax.clear();
    };
    fasta_build_strings(fasta, content.acc2tax, K, a.six, [&](const std::string &s, uint32_t t) {
        if (s.empty()) return;
        if (chars.size() + s.size() > p.max_batch_bases - 64 || starts.size() + 1 >= p.max_batch_seqs) flush();
        starts.push_back(chars.size());
        lens.push_back((uint32_t)s.size());
        tax.push_back(t);
        chars += s;
    });
    flush();
    finish_and_write(ctx, out_path, K, content, a.content.empty() && update ? a.index + "_content.txt" : a.content, kWriteAll);
    kasa_ctx_destroy(ctx);
    return 0;
}

// ------------------------------------------------------------------------------------------------ identify
struct Hit {
    size_t idx;
    float score;
    double rel;
};

// `shared` (identify_multiple): the contexts with the index loaded outlive the call — created by the first input, used by the next ones
int run_identify(const Args &a, std::vector<kasa_ctx *> *shared = nullptr)
{
    Content content = load_content(a.content);
    const size_t T = content.taxids.size();
    // index: <idx>_info.txt = count [\n 128 | 3] (main.cpp:982-991)
    uint64_t n_idx = 0;
    int bits = 64;
    bool half = false;
    {
        std::ifstream f(a.index + "_info.txt");
        if (!f) throw std::runtime_error("Index info file not found: " + a.index + "_info.txt");
        std::string tok;
        f >> n_idx;
        if (f >> tok) {
            if (tok == "128") bits = 128;
            else if (tok == "3") half = true;               // `kASA half`: 6-byte records, the upper six letters in <index>_trie
        }
    }
    const int K = bits == 128 ? 25 : 12;
    const int k_high = std::min(a.k_high, K), k_low = a.k_low, numK = k_high - k_low + 1;
    // frequencies: column for k_high (Compare.hpp:165-176)
    std::vector<uint64_t> freq_high(T, 0);
    {
        std::ifstream f(a.index + "_f.txt");
        if (!f) throw std::runtime_error("The content file or the frequency file cannot be found!");
        std::string line;
        size_t i = 0;
        while (std::getline(f, line) && i < T) {
            if (line.empty()) continue;
            auto v = split(line, '\t');
            const size_t nk = v.size() - 1;
            freq_high[i++] = std::stoull(v[1 + nk - k_high]);
        }
    }
    // The input is streamed: a batch of reads is parsed while the devices work on the earlier ones (the reference reads its input
    // through a 2048-byte window in two passes, Utilities.hpp:448-550; nothing here holds more than the batches in flight).
    ReadStream in1(a.in1.empty() ? a.input : a.in1);
    const bool paired = !a.in2.empty();
    std::unique_ptr<ReadStream> in2;
    if (paired) in2.reset(new ReadStream(a.in2));

    kasa_params p{};
    if (a.one && a.six) throw std::runtime_error("--one and --six exclude each other");     // main.cpp:618
    p.device = a.device; p.highest_k = K; p.k_high = k_high; p.k_low = k_low; p.frames = a.one ? 1 : (a.six ? 6 : 3);
    p.score_mode = a.rtt.empty() ? KASA_SCORE_NONE : (a.exact ? KASA_SCORE_EXACT : KASA_SCORE_ORDERED);
    uint8_t codon[512];
    if (!a.alphabet_file.empty() && load_codon_table(a.alphabet_file, a.alphabet_id, codon)) p.codon_table = codon;
    const uint32_t per_batch = std::max<uint32_t>(1, a.batch_reads);
    const uint32_t marker = 3 * (K - k_low);
    // a batch closes at --batch-reads reads or --batch-bases bases, whichever comes first (long reads: fewer per batch)
    const uint64_t budget_bases = a.batch_bases ? a.batch_bases : std::max<uint64_t>((uint64_t)per_batch * 320, 1ull << 24);
    std::ofstream out;
    if (!a.rtt.empty()) {
        out.open(a.rtt);
        if (!out) throw std::runtime_error("Readwise output file could not be created!");
        if (a.format == "tsv") out << "#Read number\tSpecifier from input file\tMatched taxa\tNames\tScores{relative,k-mer}\tError\n";
        else if (a.format == "json") out << "[\n";
    }
    const int frames = a.one ? 1 : (a.six ? 6 : 3);
    uint64_t n_kmers_total = 0, read_no = 0;
    // two host batches in flight per GPU (kasa_identify_submit / _wait): while the devices work on the later batches, the
    // oldest one's per-read output is formatted and written here, in input order
    // (the staging buffers are page-locked, kasa_host_alloc: only then does submit return before the copies and kernels ran)
    struct HostBatch {
        Pinned<char> chars;
        Pinned<uint64_t> starts;
        Pinned<uint32_t> lens, first;
        Pinned<float> scores;
        std::vector<std::string> names;       // per read: "<name> " (paired: both mates' names)
        std::vector<uint32_t> lengths;        // per read: the reference's "Length" (letters + one per sequence line, Utilities.hpp:521)
        uint64_t kmers_bound = 0;             // upper bound of the batch's k-mers
        kasa_batch_result res{};
        size_t n = 0;
        kasa_ctx *ctx = nullptr;
    };
    const int G = a.gpus;
    std::vector<HostBatch> hb(2 * (size_t)G);
    const size_t fmt_threads = std::max(1u, std::thread::hardware_concurrency());      // the reference formats on its worker threads too (Compare.hpp:3391)
    ReadRec pend1, pend2;
    bool have_pending = false, eof = false;
    // parses the next batch into h; false when the input is exhausted and the batch is empty
    auto fill = [&](HostBatch &h) -> bool {
        h.chars.clear(); h.starts.clear(); h.lens.clear(); h.first.clear(); h.names.clear(); h.lengths.clear();
        h.n = 0; h.kmers_bound = 0;
        uint64_t bases = 0;
        while (h.n < per_batch) {
            if (!have_pending) {
                if (!in1.next(pend1)) { eof = true; if (paired && in2->next(pend2)) throw std::runtime_error("paired-end files differ in their number of reads"); break; }
                if (paired && !in2->next(pend2)) throw std::runtime_error("paired-end files differ in their number of reads");
                have_pending = true;
            }
            const uint64_t L = pend1.seq.size() + (paired ? pend2.seq.size() : 0);
            if (h.n && bases + L > budget_bases) break;                     // the pending read opens the next batch
            if (L > budget_bases) throw std::runtime_error("a read is longer than --batch-bases (" + std::to_string(budget_bases) + "): raise it");
            have_pending = false;
            h.first.push_back((uint32_t)h.starts.size());
            std::string name = pend1.name + " ";
            uint32_t length = (uint32_t)pend1.seq.size() + pend1.lines;
            for (int m = 0; m < (paired ? 2 : 1); ++m) {
                const std::string &q = (m ? pend2 : pend1).seq;
                h.starts.push_back(h.chars.size()); h.lens.push_back((uint32_t)q.size()); h.chars.append(q.data(), q.size());
                if (q.size() + marker > (size_t)(3 * K + 1)) h.kmers_bound += (q.size() + marker - 3 * K + 1) * (a.six ? 2 : 1);
            }
            if (paired) { name += pend2.name + " "; length += (uint32_t)pend2.seq.size() + pend2.lines; }
            h.names.push_back(std::move(name)); h.lengths.push_back(length);
            bases += L;
            ++h.n;
        }
        h.first.push_back((uint32_t)h.starts.size());
        return h.n > 0;
    };
    // The first batch is parsed before the contexts exist: an input that ends inside it (the usual small run) sizes the
    // workspaces exactly, a longer one by the batch limits.
    const bool any_reads = fill(hb[0]);
    if (eof && !have_pending && !shared) {
        p.max_batch_kmers = hb[0].kmers_bound + 4096; p.max_batch_seqs = (uint32_t)hb[0].starts.size() + 1; p.max_batch_bases = hb[0].chars.size() + 4096;
    } else {
        const uint64_t seqs = (uint64_t)per_batch * (paired ? 2 : 1);
        p.max_batch_kmers = std::min<uint64_t>((a.six ? 2 : 1) * (budget_bases + seqs * marker) + 4096, (1ull << 32) - (1u << 20));
        p.max_batch_seqs = (uint32_t)seqs + 1; p.max_batch_bases = budget_bases + 4096;
    }
    // --gpus N: the index is replicated on devices device .. device + N - 1, batches are dealt round-robin (reads are
    // independent; the reference's worker threads, Compare.hpp:3221-3310, become GPUs).  One host thread drives all of them:
    // kasa_identify_submit does not wait for the device.
    std::vector<kasa_ctx *> ctxs(G, nullptr);
    if (shared && !shared->empty()) ctxs = *shared;
    else {
        std::ifstream f(a.index, std::ios::binary);
        if (!f) throw std::runtime_error("Index file not found: " + a.index);
        std::vector<char> bytes(n_idx * (half ? 6 : bits == 128 ? 20 : 12)), trie;
        f.read(bytes.data(), (std::streamsize)bytes.size());
        uint64_t n_trie = 0;
        if (half) {
            std::ifstream ti(a.index + "_trie.txt"), tf(a.index + "_trie", std::ios::binary);
            if (!ti || !tf) throw std::runtime_error("Trie file not found");                      // main.cpp:362-366
            ti >> n_trie;
            trie.resize(n_trie * 12);
            tf.read(trie.data(), (std::streamsize)trie.size());
        }
        const bool same_device = getenv("KASA_CLI_SAME_DEVICE") != nullptr;   // test hook: N contexts on one GPU
        for (int g = 0; g < G; ++g) {
            p.device = same_device ? a.device : a.device + g;
            check(kasa_ctx_create(&p, &ctxs[g]), nullptr);
            if (half) check(kasa_index_load_half(ctxs[g], bytes.data(), n_idx, trie.data(), n_trie, (uint32_t)T), ctxs[g]);
            else check(kasa_index_load(ctxs[g], bytes.data(), n_idx, bits, content.taxids.data(), (uint32_t)T), ctxs[g]);
        }
        if (shared) *shared = ctxs;
    }

    auto submit = [&](HostBatch &h, kasa_ctx *ctx) {
        h.ctx = ctx;
        kasa_read_batch in{};
        in.n_reads = (uint32_t)h.n; in.n_seqs = (uint32_t)h.starts.size(); in.format = KASA_SEQ_ASCII; in.location = KASA_MEM_HOST;
        in.bases = h.chars.data(); in.seq_start = h.starts.data(); in.seq_len = h.lens.data(); in.read_first_seq = h.first.data();
        h.res = kasa_batch_result{};
        h.scores.assign(a.rtt.empty() ? 0 : h.n * T, 0.f);
        h.res.scores = h.scores.empty() ? nullptr : h.scores.data(); h.res.scores_location = KASA_MEM_HOST;
        check(kasa_identify_submit(ctx, &in, &h.res), ctx);
    };
    auto emit = [&](HostBatch &h) {
        kasa_ctx *ctx = h.ctx;
        check(kasa_identify_wait(ctx), ctx);                  // the oldest batch of that GPU = the oldest batch overall
        const Pinned<float> &scores = h.scores;
        n_kmers_total += h.res.n_kmers;
        if (a.rtt.empty()) { read_no += h.n; return; }
        // ---- per-read output (Compare.hpp:1485-1890): formatted by all host threads, chunk by chunk, written in input order
        auto format_read = [&](size_t r, uint64_t rn, std::string &o, std::vector<Hit> &hits) {
            const std::string &name = h.names[r];
            const uint32_t length = h.lengths[r];
            const float best = best_score(length, k_high, k_low, frames);
            const float *row = scores.data() + r * T;
            size_t nh = 0;
            for (size_t t = 1; t < T; ++t)
                if (row[t] > 0.f) {
                    const double rel = row[t] / (1.0 + std::log2((double)freq_high[t] * double(length - K * 3 + 1)));
                    if (rel >= a.threshold) hits[nh++] = Hit{t, row[t], rel};
                }
            if (nh == 0) {
                if (a.format == "tsv") { o += std::to_string(rn) + "\t" + name + "\t-\t-\t-\t-\n"; }
                else if (a.format == "json") {
                    o += rn == 0 ? "{\n" : ",\n{\n";
                    o += "\t\"Read number\": " + std::to_string(rn) + ",\n\t\"Specifier from input file\": \"" + name + "\",\n\t\"Length\": " +
                         std::to_string(length) + ",\n\t\"Top hits\": [\n\t],\n\t\"Further hits\": [\n\t]\n}";
                } else if (a.format == "jsonl") {
                    o += "{ \"Read number\": " + std::to_string(rn) + ", \"Specifier from input file\": \"" + name + "\", \"Length\": " +
                         std::to_string(length) + ", \"Top hits\": [], \"Further hits\": [] }\n";
                } else {
                    o += "U\t" + name + "\t0\t";
                    o += (char)length;                        // the reference appends the length as a raw char here (Compare.hpp:1568)
                    o += "\tA:00\n";
                }
                return;
            }
            std::stable_sort(hits.begin(), hits.begin() + nh, [](const Hit &x, const Hit &y) { return x.rel > y.rel; });
            float maxv = 0.f;
            for (size_t i = 0; i < nh; ++i) maxv = std::max(maxv, hits[i].score);
            size_t top = 1;
            for (size_t i = 1; i < nh && (int64_t)i < a.beasts; ++i) { if (hits[i].score / maxv > 0.8f) ++top; else break; }
            auto hit_json = [&](const Hit &h, bool lines) {
                const char *sep = lines ? " " : "\n\t\t";
                std::string s;
                s += lines ? "{" : "\t{";
                s += std::string(sep) + "\"tax ID\": \"" + std::to_string(content.taxids[h.idx]) + "\",";
                s += std::string(sep) + "\"Name\": \"" + content.names[h.idx] + "\",";
                s += std::string(sep) + "\"k-mer Score\": "; append_double((double)h.score, s); s += ",";
                s += std::string(sep) + "\"Relative Score\": "; append_double(h.rel, s); s += ",";
                s += std::string(sep) + "\"Error\": "; append_double((double)((best - h.score) / best), s);
                s += lines ? "}" : "\n\t}";
                return s;
            };
            // entries after the top hits: up to `beasts` distinct k-mer scores (Compare.hpp:1721-1752)
            auto further = [&](auto fn) {
                float before = 0;
                for (size_t j = top, i = top; i < nh && (int64_t)j < a.beasts; ++i) {
                    fn(hits[i], i == top);
                    if (before != hits[i].score) { before = hits[i].score; ++j; }
                }
            };
            if (a.format == "jsonl") {
                o += "{ \"Read number\": " + std::to_string(rn) + ", \"Specifier from input file\": \"" + name + "\", \"Length\": " +
                     std::to_string(length) + ", \"Top hits\": [";
                for (size_t i = 0; i < top; ++i) { if (i) o += ","; o += hit_json(hits[i], true); }
                o += "], \"Further hits\": [";
                further([&](const Hit &h, bool firstOne) { if (!firstOne) o += ", "; o += hit_json(h, true); });
                o += "] }\n";
            } else if (a.format == "json") {
                o += rn == 0 ? "{\n" : ",\n{\n";
                o += "\t\"Read number\": " + std::to_string(rn) + ",\n\t\"Specifier from input file\": \"" + name + "\",\n\t\"Length\": " +
                     std::to_string(length) + ",\n\t\"Top hits\": [\n";
                for (size_t i = 0; i < top; ++i) { if (i) o += ",\n"; o += hit_json(hits[i], false); }
                o += "\n\t],\n\t\"Further hits\": [\n";
                further([&](const Hit &h, bool firstOne) { if (!firstOne) o += ",\n"; o += hit_json(h, false); });
                o += "\n\t]\n}";
            } else if (a.format == "tsv") {
                std::string s1 = std::to_string(rn) + "\t" + name + "\t", s2, s3, s4;
                float before = 0;
                for (size_t j = 0, i = 0; i < nh && (int64_t)j < a.beasts; ++i) {
                    s1 += std::to_string(content.taxids[hits[i].idx]) + ";";
                    s2 += content.names[hits[i].idx] + ";";
                    append_double(hits[i].rel, s3); s3 += ","; append_double((double)hits[i].score, s3); s3 += ";";
                    append_double((double)((best - hits[i].score) / best), s4); s4 += ";";
                    if (before != hits[i].score) { before = hits[i].score; ++j; }
                }
                for (std::string *s : {&s1, &s2, &s3, &s4}) if (!s->empty() && s->back() == ';') s->pop_back();
                if (!s2.empty()) o += s1 + "\t" + s2 + "\t" + s3 + "\t" + s4 + "\n";
            } else {
                o += "C\t" + name + "\t" + std::to_string(content.taxids[hits[0].idx]) + "\t" + std::to_string(length) + "\t";
                for (size_t i = 0; i < top; ++i) { o += std::to_string(content.taxids[hits[i].idx]) + ":"; append_double((double)hits[i].score, o); o += " "; }
                further([&](const Hit &h, bool) { o += std::to_string(content.taxids[h.idx]) + ":"; append_double((double)h.score, o); o += " "; });
                o += "\n";
            }
        };
        const size_t n_batch = h.n;
        const size_t n_thr = std::max<size_t>(1, std::min<size_t>(fmt_threads, n_batch / 256 + 1));
        std::vector<std::string> pieces(n_thr);
        auto work = [&](size_t t) {
            std::vector<Hit> hits(T);
            const size_t r0 = n_batch * t / n_thr, r1 = n_batch * (t + 1) / n_thr;
            pieces[t].reserve((r1 - r0) * 256);
            for (size_t r = r0; r < r1; ++r) format_read(r, read_no + r, pieces[t], hits);
        };
        if (n_thr == 1) work(0);
        else {
            std::vector<std::thread> pool;
            for (size_t t = 0; t < n_thr; ++t) pool.emplace_back(work, t);
            for (auto &th : pool) th.join();
        }
        for (const std::string &piece : pieces) out << piece;
        read_no += n_batch;
    };
    // ring of 2 G host batches: batch nb runs on GPU nb % G; before its slot is reused, the batch that held it (2 G batches
    // ago, the oldest in flight) is emitted
    const size_t ring = hb.size();
    size_t nb = 0, emitted = 0;
    for (bool more = any_reads; more; ++nb) {
        submit(hb[nb % ring], ctxs[nb % (size_t)G]);
        if (nb + 1 - emitted == ring) emit(hb[emitted++ % ring]);       // frees the slot the next batch is parsed into
        more = fill(hb[(nb + 1) % ring]);
    }
    while (emitted < nb) emit(hb[emitted++ % ring]);
    if (!a.rtt.empty() && a.format == "json") out << "\n]";
    out.close();

    // ---- profile (Compare.hpp:3466-3665)
    std::vector<double> ca(numK * T, 0.0);
    std::vector<uint64_t> cu(numK * T, 0);
    for (kasa_ctx *c : ctxs) {                                // the per-thread reduction of Compare.hpp:3445-3454, over GPUs
        std::vector<double> ca1(numK * T);
        std::vector<uint64_t> cu1(numK * T);
        check(kasa_identify_finish(c, ca1.data(), cu1.data(), nullptr), c);
        for (size_t i = 0; i < ca.size(); ++i) { ca[i] += ca1[i]; cu[i] += cu1[i]; }
    }
    std::vector<uint64_t> sum_u(numK, 0);
    std::vector<double> sum_n(numK, 0.0);
    for (size_t t = 1; t < T; ++t)
        for (int ik = 0; ik < numK; ++ik) { sum_u[ik] += cu[ik * T + t]; sum_n[ik] += ca[ik * T + t]; }
    if (!a.profile.empty()) {
        std::vector<size_t> order;
        for (size_t t = 0; t < T; ++t) order.push_back(t);
        std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) {
            for (int ik = 0; ik < numK; ++ik) {
                const uint64_t ux = x ? cu[ik * T + x] : 0, uy = y ? cu[ik * T + y] : 0;
                if (ux != uy) return ux > uy;
            }
            return false;
        });
        std::vector<uint64_t> garbage(numK, 0);
        for (int i = k_high - k_low, j = 0; i > 0; --i, ++j) garbage[j] = read_no * (uint64_t)frames * i;
        std::ofstream tf(a.profile);
        if (!tf) throw std::runtime_error("Profile file couldn't be opened for writing!");
        std::ostringstream body;
        tf << "#taxID,Name";
        for (const char *h : {"Unique counts k=", "Unique rel. freq. k=", "Non-unique counts k=", "Non-unique rel. freq. k=", "Overall rel. freq. k=",
                              "Overall unique rel. freq. k="})
            for (int ik = 0; ik < numK; ++ik) tf << "," << h << k_high - ik;
        tf << "\n";
        std::vector<double> sum_id(numK, 0), sum_uid(numK, 0);
        for (size_t t : order) {
            if (t == 0 || !(ca[(numK - 1) * T + t] > 0)) continue;
            body << content.taxids[t] << "," << content.names[t];
            for (int ik = 0; ik < numK; ++ik) body << "," << cu[ik * T + t];
            for (int ik = 0; ik < numK; ++ik) { if (cu[ik * T + t] == 0) body << "," << 0.0; else body << "," << (double)cu[ik * T + t] / sum_u[ik]; }
            for (int ik = 0; ik < numK; ++ik) body << "," << ca[ik * T + t];
            for (int ik = 0; ik < numK; ++ik) { if (ca[ik * T + t] == 0) body << "," << 0.0; else body << "," << ca[ik * T + t] / sum_n[ik]; }
            for (int ik = 0; ik < numK; ++ik) { sum_id[ik] += ca[ik * T + t]; body << "," << ca[ik * T + t] / (double)(n_kmers_total - garbage[ik]); }
            for (int ik = 0; ik < numK; ++ik) { sum_uid[ik] += (double)cu[ik * T + t]; body << "," << (double)cu[ik * T + t] / (double)(n_kmers_total - garbage[ik]); }
            body << "\n";
        }
        tf << "0,not identified";
        for (int i = 0; i < numK * 4; ++i) tf << "," << 0.0;
        for (int ik = 0; ik < numK; ++ik) tf << "," << ((double)n_kmers_total - (double)garbage[ik] - sum_id[ik]) / ((double)n_kmers_total - (double)garbage[ik]);
        for (int ik = 0; ik < numK; ++ik) tf << "," << ((double)n_kmers_total - (double)garbage[ik] - sum_uid[ik]) / ((double)n_kmers_total - (double)garbage[ik]);
        tf << "\n" << body.str();
    }
    if (a.verbose) {
        double all = 0;
        for (size_t t = 1; t < T; ++t) all += ca[(numK - 1) * T + t];
        std::cout << "OUT: Number of k-mers in input: " << n_kmers_total << " of which " << all / n_kmers_total * 100. << " % were identified." << std::endl;
        std::cout << "OUT: Number of uniques:";
        for (int ik = 0; ik < numK; ++ik) std::cout << " " << sum_u[ik];
        std::cout << std::endl;
    }
    if (!shared) for (kasa_ctx *c : ctxs) kasa_ctx_destroy(c);
    return 0;
}

// identify_multiple (main.cpp:1118-1334): every file of the folder -i is identified against the index, which is loaded once; the
// outputs are <-q prefix><file name without its last extension><.json|.jsonl|.tsv|.ktsv> and <-p prefix><name>.csv.  The reference
// runs the files on its worker threads; here they pass through the same device contexts one after the other, each pipelined
// batch by batch like a single identify.
int run_identify_multiple(const Args &a)
{
    if (!a.in1.empty() || !a.in2.empty()) throw std::runtime_error("No paired-end allowed in this mode!");      // main.cpp:1129
    std::vector<std::string> files = list_directory(a.input);
    if (files.empty()) throw std::runtime_error("No files found in " + a.input);
    const std::string ending = a.format == "json" ? ".json" : a.format == "jsonl" ? ".jsonl" : a.format == "tsv" ? ".tsv" : ".ktsv";
    std::vector<kasa_ctx *> ctxs;
    int rc = 0;
    for (const std::string &name : files) {
        Args b = a;
        b.mode = "identify";
        b.input = a.input + (a.input.empty() || a.input.back() == '/' ? "" : "/") + name;
        const size_t dot = name.rfind('.');
        const std::string stem = dot == std::string::npos ? name : name.substr(0, dot);
        b.rtt = a.rtt.empty() ? "" : a.rtt + stem + ending;
        b.profile = a.profile.empty() ? "" : a.profile + stem + ".csv";
        rc |= run_identify(b, &ctxs);
    }
    for (kasa_ctx *c : ctxs) kasa_ctx_destroy(c);
    return rc;
}

}  // namespace

int main(int argc, char **argv)
{
    try {
        const auto t0 = std::chrono::steady_clock::now();
        Args a = parse(argc, argv);
        int rc;
        if (a.show) {                                       // host-only: the parameters as read from --parameters
            std::cout << "mode=" << a.mode << "\nindex=" << a.index << "\ncontent=" << a.content << "\ninput=" << a.input << "\nin1=" << a.in1 << "\nin2=" << a.in2
                      << "\nprofile=" << a.profile << "\nrtt=" << a.rtt << "\nformat=" << a.format << "\nk=" << a.k_high << " " << a.k_low << "\nhighest_k=" << a.highest_k
                      << "\nframes=" << (a.one ? 1 : a.six ? 6 : 3) << "\nbeasts=" << a.beasts << "\nthreshold=" << a.threshold << "\nalphabet=" << a.alphabet_file << " "
                      << a.alphabet_id << "\nout=" << a.out << "\nfirst=" << a.first_index << "\nsecond=" << a.second_index << "\ndeleted=" << a.deleted
                      << "\nverbose=" << a.verbose << std::endl;
            return 0;
        }
        if (a.mode == "host-build-strings") {              // host-only (no device): the strings `build` would push
            Content c = load_content(a.content);
            fasta_build_strings(slurp(a.input), c.acc2tax, a.highest_k, a.six,
                                [](const std::string &s, uint32_t t) { std::cout << t << "\t" << s << "\n"; });
            return 0;
        }
        if (a.mode == "host-reads") {                      // host-only: parsed reads (name, +lines, sequence)
            for (const ReadRec &r : parse_reads(slurp(a.input))) std::cout << r.name << "\t" << r.lines << "\t" << r.seq << "\n";
            return 0;
        }
        if (a.mode == "build" || a.mode == "update") rc = run_build(a);
        else if (a.mode == "identify") rc = run_identify(a);
        else if (a.mode == "identify_multiple") rc = run_identify_multiple(a);
        else if (a.mode == "half") rc = run_half(a);
        else if (a.mode == "redundancy") rc = run_redundancy(a);
        else if (a.mode == "trie" || a.mode == "getFrequency" || a.mode == "delete" || a.mode == "merge") rc = run_index_mode(a);
        else if (a.mode == "-h" || a.mode == "--help") usage();
        else throw std::runtime_error("mode not supported by kasa_b200: " + a.mode + " (build, identify)");
        std::cout << "OUT: Time: " << std::chrono::duration_cast<std::chrono::seconds>(std::chrono::steady_clock::now() - t0).count() << " s" << std::endl;
        return rc;
    } catch (const std::exception &e) {
        std::cerr << "ERROR: " << e.what() << std::endl;      // main.cpp:1717-1720
        return 1;
    }
}
