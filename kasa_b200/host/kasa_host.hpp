// kasa_host.hpp — host side of the drop-in: the parts of kASA that stay on the CPU (SURVEY.md §8b "outer"
// boundary): FASTA/FASTQ(.gz) reading, content file, index/trie/frequency file layouts, per-read scoring and the
// profile / JSON / TSV / Kraken writers.  All compute goes through the C ABI of include/kasa_b200.h.
// Written from the reference's observable behaviour; file:line cites /root/reference/source.
#pragma once
#include <zlib.h>

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace kasa_host {

// ------------------------------------------------------------------------------------------------ files
// whole file into memory, transparently inflating .gz (zlib's gz* API reads plain files too)
inline std::string slurp(const std::string &path)
{
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("File couldn't be opened: " + path);
    gzbuffer(f, 1 << 20);
    std::string out;
    std::vector<char> buf(1 << 22);
    int n;
    while ((n = gzread(f, buf.data(), (unsigned)buf.size())) > 0) out.append(buf.data(), (size_t)n);
    gzclose(f);
    if (n < 0) throw std::runtime_error("Error while reading " + path);
    return out;
}

inline std::vector<std::string> split(const std::string &s, char d)
{
    std::vector<std::string> out;
    std::string cur;
    std::istringstream ss(s);
    while (std::getline(ss, cur, d)) out.push_back(cur);
    return out;
}

// ------------------------------------------------------------------------------------------------ content file
struct Content {
    std::vector<std::string> names;      // index 0 = "non_unique" (Compare.hpp:124-126)
    std::vector<uint32_t> taxids;        // taxid_of_idx
    std::unordered_map<std::string, uint32_t> acc2tax;
};

// Compare.hpp:111-150 (identify) and Read.hpp:2955-3008 (build): name \t taxid \t taxids \t accessions(;)
inline Content load_content(const std::string &path)
{
    std::ifstream f(path);
    if (!f) throw std::runtime_error("The content file cannot be found!");
    Content c;
    c.names.push_back("non_unique");
    c.taxids.push_back(0);
    std::string line;
    while (std::getline(f, line)) {
        if (line.empty()) continue;
        auto p = split(line, '\t');
        if (p.size() < 4) throw std::runtime_error("Content file contains less than 4 columns, it may be damaged... The faulty line was: " + line + "\n");
        std::string name = p[0];
        name.erase(std::remove(name.begin(), name.end(), ','), name.end());
        c.names.push_back(name);
        const uint32_t tax = (uint32_t)std::stoul(p.size() >= 5 ? p[4] : p[1]);
        c.taxids.push_back(tax);
        for (const auto &a : split(p[3], ';'))
            if (!a.empty()) c.acc2tax[a] = tax;
    }
    return c;
}

// ------------------------------------------------------------------------------------------------ numbers
// Shortest round-trip text of a double in the layout of the reference's dtoa_milo (dToStr.h:427): plain
// decimals for exponents in (-7, 21], "d.ddde-N" otherwise, integers with a trailing ".0".
inline void append_double(double v, std::string &out)
{
    if (v == 0) { out += "0.0"; return; }
    if (v < 0) { out += '-'; v = -v; }
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::scientific);
    std::string s(buf, r.ptr);                       // d.ddddde[+-]XX
    const size_t epos = s.find('e');
    std::string digits = s.substr(0, epos);
    digits.erase(std::remove(digits.begin(), digits.end(), '.'), digits.end());
    const int exp10 = std::stoi(s.substr(epos + 1));
    const int len = (int)digits.size(), kk = exp10 + 1;     // decimal point position
    if (len <= kk && kk <= 21) { out += digits; out.append((size_t)(kk - len), '0'); out += ".0"; }
    else if (0 < kk && kk <= 21) { out.append(digits, 0, (size_t)kk); out += '.'; out.append(digits, (size_t)kk, std::string::npos); }
    else if (-6 < kk && kk <= 0) { out += "0."; out.append((size_t)(-kk), '0'); out += digits; }
    else {
        out += digits[0];
        if (len > 1) { out += '.'; out.append(digits, 1, std::string::npos); }
        out += 'e';
        out += std::to_string(kk - 1);
    }
}

// ------------------------------------------------------------------------------------------------ sequences
struct ReadRec {
    std::string name;      // header without its first character (Read.hpp:711-714)
    std::string seq;       // raw letters
    uint32_t lines = 0;    // number of sequence lines: the reference's "Length" counts their newlines (Utilities.hpp:521)
};

// FASTQ (4-line or multi-line) / FASTA records of a whole file
inline std::vector<ReadRec> parse_reads(const std::string &data)
{
    std::vector<ReadRec> out;
    size_t pos = 0;
    const size_t n = data.size();
    auto next_line = [&](std::string &line) -> bool {
        if (pos >= n) return false;
        size_t e = data.find('\n', pos);
        if (e == std::string::npos) e = n;
        line.assign(data, pos, e - pos);
        if (!line.empty() && line.back() == '\r') line.pop_back();
        pos = e + 1;
        return true;
    };
    std::string line;
    bool have = next_line(line);
    while (have) {
        if (line.empty()) { have = next_line(line); continue; }
        if (line[0] == '>') {
            ReadRec r;
            r.name = line.substr(1);
            while ((have = next_line(line)) && (line.empty() || line[0] != '>')) {
                if (line.empty()) continue;
                r.seq += line; ++r.lines;
            }
            out.push_back(std::move(r));
        } else if (line[0] == '@') {
            ReadRec r;
            r.name = line.substr(1);
            while ((have = next_line(line)) && (line.empty() || line[0] != '+')) {
                if (line.empty()) continue;
                r.seq += line; ++r.lines;
            }
            size_t q = 0;                    // quality: as many characters as bases
            while (q < r.seq.size() && (have = next_line(line))) q += line.size();
            have = next_line(line);
            out.push_back(std::move(r));
        } else {
            throw std::runtime_error("Input is neither FASTA nor FASTQ.");
        }
    }
    return out;
}

// The same records one at a time from a (possibly gzip-compressed) file: identify streams its input batch by batch instead of holding
// the whole file (the reference reads it through FileReader::getChunk, Utilities.hpp:448-550, in two passes).
class ReadStream {
public:
    explicit ReadStream(const std::string &path) : f_(gzopen(path.c_str(), "rb")), buf_(1 << 22)
    {
        if (!f_) throw std::runtime_error("File couldn't be opened: " + path);
        gzbuffer(f_, 1 << 20);
        have_ = next_line(line_);
    }
    ReadStream(const ReadStream &) = delete;
    ReadStream &operator=(const ReadStream &) = delete;
    ~ReadStream() { if (f_) gzclose(f_); }

    // next record into r; false at the end of the file
    bool next(ReadRec &r)
    {
        while (have_ && line_.empty()) have_ = next_line(line_);
        if (!have_) return false;
        r.name.clear(); r.seq.clear(); r.lines = 0;
        if (line_[0] == '>') {
            r.name = line_.substr(1);
            while ((have_ = next_line(line_)) && (line_.empty() || line_[0] != '>')) {
                if (line_.empty()) continue;
                r.seq += line_; ++r.lines;
            }
        } else if (line_[0] == '@') {
            r.name = line_.substr(1);
            while ((have_ = next_line(line_)) && (line_.empty() || line_[0] != '+')) {
                if (line_.empty()) continue;
                r.seq += line_; ++r.lines;
            }
            size_t q = 0;                    // quality: as many characters as bases
            while (q < r.seq.size() && (have_ = next_line(line_))) q += line_.size();
            have_ = next_line(line_);
        } else {
            throw std::runtime_error("Input is neither FASTA nor FASTQ.");
        }
        return true;
    }

private:
    bool fill()
    {
        const int n = gzread(f_, buf_.data(), (unsigned)buf_.size());
        if (n < 0) throw std::runtime_error("Error while reading the input");
        pos_ = 0; end_ = (size_t)n;
        return n > 0;
    }
    bool next_line(std::string &line)
    {
        line.clear();
        bool any = false;
        for (;;) {
            if (pos_ == end_ && !fill()) break;
            any = true;
            const char *b = buf_.data() + pos_;
            const char *nl = static_cast<const char *>(std::memchr(b, '\n', end_ - pos_));
            if (nl) { line.append(b, (size_t)(nl - b)); pos_ += (size_t)(nl - b) + 1; break; }
            line.append(b, end_ - pos_);
            pos_ = end_;
        }
        if (!line.empty() && line.back() == '\r') line.pop_back();
        return any;
    }
    gzFile f_;
    std::vector<char> buf_;
    size_t pos_ = 0, end_ = 0;
    std::string line_;
    bool have_ = false;
};

// ------------------------------------------------------------------------------------------------ build: FASTA pieces
// The reference reads the FASTA through a 2048-byte buffer; a chunk ends at a newline or at the buffer end
// (Utilities::FileReader::getChunk, Utilities.hpp:448-550).  Piece boundaries, and with --six the places where the
// reverse strand gets a marker, depend on this chunking (Read.hpp:2306-2578), so it is reproduced here.
class ChunkReader {
public:
    explicit ChunkReader(const std::string &d) : data_(d) {}
    // returns false at end of data; `complete` = chunk ended at a newline
    bool next(std::string &chunk, bool &complete)
    {
        chunk.clear();
        const size_t n = data_.size();
        if (pos_ >= n) { eof_ = true; complete = false; return false; }
        const size_t bufend = (pos_ / 2048 + 1) * 2048;
        const size_t lim = std::min(bufend, n);
        const void *nl = std::memchr(data_.data() + pos_, '\n', lim - pos_);
        if (nl) {
            const size_t e = (const char *)nl - data_.data();
            chunk.assign(data_, pos_, e - pos_); pos_ = e + 1; complete = true;
        } else if (bufend <= n) {
            chunk.assign(data_, pos_, bufend - pos_); pos_ = bufend; complete = false;
        } else {
            chunk.assign(data_, pos_, n - pos_); pos_ = n; complete = true;      // sentinel newline behind the last partial buffer
        }
        return true;
    }
    bool eof() const { return eof_; }
private:
    const std::string &data_;
    size_t pos_ = 0;
    bool eof_ = false;
};

inline std::string accession_of(const std::string &header)
{
    const std::string first = header.substr(0, header.find(' '));
    for (const auto &part : split(first, '|'))
        if (part.find('.') != std::string::npos) return part;
    return "";
}

inline void clean_dna(std::string &s)          // Read.hpp:2504-2520: anything but ACGTacgt becomes 'Z'
{
    for (char &c : s) {
        if (c == '\t' || c == ' ') throw std::runtime_error("Spaces or tabs inside reference, please check your input.");
        if (c != 'A' && c != 'C' && c != 'G' && c != 'T' && c != 'a' && c != 'c' && c != 'g' && c != 't') c = 'Z';
    }
}

inline std::string revcomp(const std::string &s)     // kASA.hpp:214-221, table kASA.hpp:54
{
    static const char rc[8] = {'T', 'G', 'A', 'C', 'X', 'Z', '?', '?'};
    std::string o(s.size(), ' ');
    for (size_t j = 0; j < s.size(); ++j) o[s.size() - 1 - j] = rc[(s[j] >> 1) & 7];
    return o;
}

// Calls emit(string, taxid) for every string the reference hands to dnaTokMers.
template <class Emit>
void fasta_build_strings(const std::string &data, const std::unordered_map<std::string, uint32_t> &acc2tax, int K, bool six, Emit emit)
{
    const std::string marker((size_t)(3 * (K - 1)), 'X');            // Read.hpp:2323-2333 with _iLowestK = 1
    ChunkReader rd(data);
    uint32_t idx = 0, next_idx = 0;
    bool found = false, complete = false;
    std::string chunk;
    auto lookup = [&](const std::string &hdr, uint32_t &dst) -> bool {
        auto it = acc2tax.find(accession_of(hdr));
        if (it == acc2tax.end()) it = acc2tax.find(hdr);
        if (it == acc2tax.end()) return false;
        dst = it->second;
        return true;
    };
    while (!rd.eof()) {                                               // header of the first entry
        if (!rd.next(chunk, complete)) break;
        if (chunk.empty()) continue;
        if (chunk[0] != '>') throw std::runtime_error("No > found in input.");
        found = lookup(chunk.substr(1), idx);
        break;
    }
    bool add_marker = false, next_flag = false, add_rc_marker = true;
    std::string overhang, dna;
    while (!rd.eof() || !overhang.empty()) {
        if (next_flag) { add_marker = false; add_rc_marker = true; }
        bool create_overhang = true;
        for (int i = 0; i < 20;) {
            const bool got = rd.next(chunk, complete);
            if (!got || chunk.empty()) {
                if (!rd.eof()) break;
                create_overhang = false; add_marker = true;
                break;
            }
            if (chunk[0] == '>') {
                std::string hdr = chunk.substr(1), more;
                while (!complete && rd.next(more, complete)) hdr += more;
                uint32_t t;
                if (lookup(hdr, t)) { next_idx = t; found = true; next_flag = true; }
                else found = false;
                create_overhang = false; add_marker = true; add_rc_marker = true;
                break;
            }
            if (found) { dna += chunk; ++i; }                          // lines of unknown accessions are skipped uncounted
        }
        dna = overhang + dna;
        if (dna.size() < (size_t)(3 * K + 2) && create_overhang) { overhang = dna; dna.clear(); continue; }
        clean_dna(dna);
        std::string rc = six ? revcomp(dna) : std::string();
        if (add_rc_marker && !rc.empty()) { rc += marker; add_rc_marker = false; }
        if (add_marker && !dna.empty()) { dna += marker; add_marker = false; }
        emit(dna, idx);
        if (six) emit(rc, idx);
        if (next_flag) { idx = next_idx; next_flag = false; }
        overhang = create_overhang ? dna.substr(dna.size() + 1 - 3 * K) : std::string();
        dna.clear();
    }
}

// ------------------------------------------------------------------------------------------------ index files
// STXXL vector files are whole blocks (MetaHeader.h:137-138): 2 101 248 B for 12-byte records, 2 048 000 B for 20-byte ones
inline void write_padded(const std::string &path, const void *data, size_t bytes, size_t block)
{
    std::ofstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("File couldn't be created: " + path);
    f.write((const char *)data, (std::streamsize)bytes);
    const size_t total = std::max<size_t>(1, (bytes + block - 1) / block) * block;
    std::vector<char> zeros(1 << 20, 0);
    for (size_t left = total - bytes; left;) {
        const size_t n = std::min(left, zeros.size());
        f.write(zeros.data(), (std::streamsize)n);
        left -= n;
    }
}

// Compare.hpp:1452-1481 calculateBestScore (DNA)
inline float best_score(uint32_t length, int k_high, int k_low, int frames)
{
    float best = 0.f;
    for (int k = k_low; k <= k_high; ++k) {
        const float w = (float)(k * k) / 625.f;
        if (frames == 1) best += ((size_t)length / 3 - k + 1) * w;          // --one (Compare.hpp:1456-1460)
        else if (frames == 6) best += 2 * ((size_t)length - k * 3 + 1) * w;
        else best += ((size_t)length - k * 3 + 1) * w;
    }
    return best;
}

}  // namespace kasa_host
