// fastx.hpp — FASTA/FASTQ (+gzip via zlib) reader with needletail 0.5.1's seq()/id() semantics
// as sylph uses them (src/sketch.rs:488,557,906): id = whole header line without the marker,
// seq = sequence with line endings stripped (multi-line FASTA joined), no case / alphabet
// normalisation. Records are appended to a flat base buffer + offsets, the layout the C ABI takes.
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace host {

struct FlatRecords {
    std::vector<uint8_t> bases;       // all sequences concatenated
    std::vector<uint64_t> offsets{0}; // n_records + 1
    std::vector<std::string> ids;     // header lines (only kept when want_ids)
    size_t n() const { return offsets.size() - 1; }
};

class LineReader {
  public:
    explicit LineReader(const std::string &path) : f_(gzopen(path.c_str(), "rb")) {
        if (f_) gzbuffer(f_, 1 << 20);
    }
    ~LineReader() { if (f_) gzclose(f_); }
    bool ok() const { return f_ != nullptr; }
    // next line without its terminator ("\n" or "\r\n"); false at EOF
    bool next(std::string &line) {
        line.clear();
        bool got = false;
        for (;;) {
            if (pos_ == len_) {
                int n = gzread(f_, buf_, sizeof buf_);
                if (n <= 0) {
                    if (got && !line.empty() && line.back() == '\r') line.pop_back();
                    return got;
                }
                len_ = (size_t)n;
                pos_ = 0;
            }
            got = true;
            const char *b = buf_ + pos_;
            const char *nl = (const char *)memchr(b, '\n', len_ - pos_);
            if (nl) {
                line.append(b, (size_t)(nl - b));
                pos_ += (size_t)(nl - b) + 1;
                if (!line.empty() && line.back() == '\r') line.pop_back();
                return true;
            }
            line.append(b, len_ - pos_);
            pos_ = len_;
        }
    }
  private:
    gzFile f_;
    char buf_[1 << 16];
    size_t pos_ = 0, len_ = 0;
};

// Appends every record of `path` to out. Returns false if the file cannot be opened or is not
// FASTA/FASTQ ("... is not a valid fasta/fastq file; skipping", src/sketch.rs:560-562).
inline bool read_fastx(const std::string &path, FlatRecords &out, bool want_ids, std::string *first_id = nullptr) {
    LineReader lr(path);
    if (!lr.ok()) return false;
    std::string line;
    if (!lr.next(line)) return false;
    while (line.empty()) if (!lr.next(line)) return false;
    bool first = true;
    if (line[0] == '>') {
        bool more = true;
        while (more) {
            if (line.empty() || line[0] != '>') return false;
            if (first && first_id) *first_id = line.substr(1);
            if (want_ids) out.ids.push_back(line.substr(1));
            first = false;
            more = false;
            while (lr.next(line)) {
                if (!line.empty() && line[0] == '>') { more = true; break; }
                out.bases.insert(out.bases.end(), line.begin(), line.end());
            }
            out.offsets.push_back(out.bases.size());
        }
        return true;
    }
    if (line[0] == '@') {
        for (;;) {
            if (line.empty() || line[0] != '@') return false;
            if (first && first_id) *first_id = line.substr(1);
            if (want_ids) out.ids.push_back(line.substr(1));
            first = false;
            std::string seq, plus, qual;
            // a record without its sequence, '+' or quality line is an invalid file (needletail reports an
            // error for the record; sylph then drops the file, src/sketch.rs:909-915)
            if (!lr.next(seq) || !lr.next(plus) || plus.empty() || plus[0] != '+' || !lr.next(qual)) return false;
            out.bases.insert(out.bases.end(), seq.begin(), seq.end());
            out.offsets.push_back(out.bases.size());
            do { if (!lr.next(line)) return true; } while (line.empty());
        }
    }
    return false;
}

}  // namespace host
