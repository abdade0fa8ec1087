// swb_io.cpp -- real sequence input for the fill (host only): FASTA, UCSC .2bit and a batch manifest.
// Replaces generate() (omp_smithW.c:489-519), which is the reference's only source of sequences
// (SURVEY 8(f)2).  Sequences come back as plain byte strings over the file's alphabet, upper-cased,
// exactly what swb_fill_async / swb_fill_multi / swb_fill_pairs_async take as `a` and `b`.
#include "../../include/swb200.h"

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

bool slurp(const char* path, std::string& out)
{
    FILE* f = std::fopen(path, "rb");
    if (!f) return false;
    char buf[1 << 16];
    size_t k;
    while ((k = std::fread(buf, 1, sizeof buf, f)) > 0) out.append(buf, k);
    std::fclose(f);
    return true;
}

struct Record { std::string name, seq; };

uint32_t rd32(const std::string& d, size_t off, bool swap)
{
    uint32_t v;
    std::memcpy(&v, d.data() + off, 4);
    return swap ? __builtin_bswap32(v) : v;
}

// UCSC .2bit: signature 0x1A412743, version 0, sequence count, reserved; index of (name, offset); per record
// dnaSize, N blocks, mask blocks, reserved, packed DNA (T 00, C 01, A 10, G 11; first base in the top bits)
bool parse_2bit(const std::string& d, std::vector<Record>& recs)
{
    if (d.size() < 16) return false;
    bool swap = false;
    uint32_t sig = rd32(d, 0, false);
    if (sig == 0x4327411Au) swap = true; else if (sig != 0x1A412743u) return false;
    if (rd32(d, 4, swap) != 0) return false;
    const uint32_t count = rd32(d, 8, swap);
    size_t off = 16;
    std::vector<std::pair<std::string, uint32_t>> index;
    for (uint32_t k = 0; k < count; ++k) {
        if (off + 1 > d.size()) return false;
        const size_t nl = (unsigned char)d[off];
        if (off + 1 + nl + 4 > d.size()) return false;
        index.emplace_back(d.substr(off + 1, nl), rd32(d, off + 1 + nl, swap));
        off += 1 + nl + 4;
    }
    static const char kBase[4] = {'T', 'C', 'A', 'G'};
    for (auto& ix : index) {
        size_t p = ix.second;
        if (p + 8 > d.size()) return false;
        const uint32_t size = rd32(d, p, swap), nblocks = rd32(d, p + 4, swap);
        p += 8;
        if (p + 8ull * nblocks + 4 > d.size()) return false;
        std::vector<uint32_t> nstart(nblocks), nsize(nblocks);
        for (uint32_t k = 0; k < nblocks; ++k) nstart[k] = rd32(d, p + 4ull * k, swap);
        for (uint32_t k = 0; k < nblocks; ++k) nsize[k] = rd32(d, p + 4ull * (nblocks + k), swap);
        p += 8ull * nblocks;
        const uint32_t mblocks = rd32(d, p, swap);
        p += 4 + 8ull * mblocks + 4;                       // mask blocks (soft masking: ignored, output is upper case), reserved
        if (p + (size + 3ull) / 4 > d.size()) return false;
        Record r;
        r.name = ix.first;
        r.seq.resize(size);
        for (uint32_t k = 0; k < size; ++k)
            r.seq[k] = kBase[((unsigned char)d[p + k / 4] >> (6 - 2 * (k & 3))) & 3];
        for (uint32_t k = 0; k < nblocks; ++k)
            for (uint64_t x = nstart[k]; x < (uint64_t)nstart[k] + nsize[k] && x < size; ++x) r.seq[x] = 'N';
        recs.push_back(std::move(r));
    }
    return true;
}

// FASTA: '>' header lines, everything else is sequence; white space is dropped, letters are upper-cased.
// A file without any header is one anonymous record (plain text sequence).
void parse_fasta(const std::string& d, std::vector<Record>& recs)
{
    size_t p = 0;
    bool open = false;
    while (p < d.size()) {
        size_t e = d.find('\n', p);
        if (e == std::string::npos) e = d.size();
        if (d[p] == '>' || d[p] == ';') {
            if (d[p] == '>') {
                Record r;
                size_t q = p + 1;
                while (q < e && !std::isspace((unsigned char)d[q])) ++q;
                r.name = d.substr(p + 1, q - p - 1);
                recs.push_back(std::move(r));
                open = true;
            }
        } else {
            if (!open) { recs.emplace_back(); open = true; }
            std::string& s = recs.back().seq;
            for (size_t q = p; q < e; ++q) {
                const unsigned char c = (unsigned char)d[q];
                if (!std::isspace(c)) s.push_back((char)std::toupper(c));
            }
        }
        p = e + 1;
    }
}

bool load_records(const char* path, std::vector<Record>& recs)
{
    std::string d;
    if (!slurp(path, d)) return false;
    if (d.size() >= 4) {
        const uint32_t sig = rd32(d, 0, false);
        if (sig == 0x1A412743u || sig == 0x4327411Au) return parse_2bit(d, recs);
    }
    parse_fasta(d, recs);
    return true;
}

char* dup_seq(const std::string& s)
{
    char* p = (char*)std::malloc(s.size() + 1);
    if (p) { std::memcpy(p, s.data(), s.size()); p[s.size()] = 0; }
    return p;
}

}  // namespace

struct swb_manifest {
    std::vector<std::string> a, b;       // the pairs' sequences
    std::vector<std::string> label;
};

extern "C" {

int swb_seq_count(const char* path, int64_t* nrecords)
{
    if (!path || !nrecords) return SWB_ERR_ARG;
    std::vector<Record> recs;
    if (!load_records(path, recs)) return SWB_ERR_IO;
    *nrecords = (int64_t)recs.size();
    return SWB_OK;
}

int swb_seq_read(const char* path, int64_t record, char** seq, int64_t* len, char* name, size_t name_cap)
{
    if (!path || !seq || !len || record < 0) return SWB_ERR_ARG;
    std::vector<Record> recs;
    if (!load_records(path, recs)) return SWB_ERR_IO;
    if ((size_t)record >= recs.size()) return SWB_ERR_ARG;
    const Record& r = recs[(size_t)record];
    *seq = dup_seq(r.seq);
    if (!*seq) return SWB_ERR_NOMEM;
    *len = (int64_t)r.seq.size();
    if (name && name_cap) std::snprintf(name, name_cap, "%s", r.name.c_str());
    return SWB_OK;
}

void swb_seq_free(char* seq) { std::free(seq); }

// manifest: one pair per line, "<fileA>[:record] <fileB>[:record]" (record = 0-based index in the file, default 0;
// '#' starts a comment; relative paths are relative to the manifest's directory)
int swb_manifest_load(const char* path, swb_manifest** out)
{
    if (!path || !out) return SWB_ERR_ARG;
    std::string d;
    if (!slurp(path, d)) return SWB_ERR_IO;
    std::string dir(path);
    const size_t slash = dir.find_last_of('/');
    dir = (slash == std::string::npos) ? std::string() : dir.substr(0, slash + 1);
    swb_manifest* m = new swb_manifest();
    std::vector<std::pair<std::string, std::vector<Record>>> cache;
    auto fetch = [&](const std::string& spec, std::string& seq) -> bool {
        std::string file = spec; long rec = 0;
        const size_t colon = spec.find_last_of(':');
        if (colon != std::string::npos && colon + 1 < spec.size() &&
            spec.find_first_not_of("0123456789", colon + 1) == std::string::npos) {
            file = spec.substr(0, colon); rec = std::atol(spec.c_str() + colon + 1);
        }
        if (!file.empty() && file[0] != '/') file = dir + file;
        for (auto& c : cache)
            if (c.first == file) { if ((size_t)rec >= c.second.size()) return false; seq = c.second[(size_t)rec].seq; return true; }
        cache.emplace_back(file, std::vector<Record>());
        if (!load_records(file.c_str(), cache.back().second)) return false;
        if ((size_t)rec >= cache.back().second.size()) return false;
        seq = cache.back().second[(size_t)rec].seq;
        return true;
    };
    size_t p = 0;
    int rc = SWB_OK;
    while (p < d.size() && rc == SWB_OK) {
        size_t e = d.find('\n', p);
        if (e == std::string::npos) e = d.size();
        std::string line = d.substr(p, e - p);
        p = e + 1;
        const size_t hash = line.find('#');
        if (hash != std::string::npos) line.resize(hash);
        char fa[4096], fb[4096];
        const int got = std::sscanf(line.c_str(), "%4095s %4095s", fa, fb);
        if (got <= 0) continue;
        if (got != 2) { rc = SWB_ERR_IO; break; }
        std::string sa, sb;
        if (!fetch(fa, sa) || !fetch(fb, sb) || sa.empty() || sb.empty()) { rc = SWB_ERR_IO; break; }
        m->a.push_back(std::move(sa)); m->b.push_back(std::move(sb));
        m->label.push_back(std::string(fa) + " " + fb);
    }
    if (rc != SWB_OK) { delete m; return rc; }
    *out = m;
    return SWB_OK;
}

int64_t swb_manifest_pairs(const swb_manifest* m) { return m ? (int64_t)m->a.size() : 0; }

int swb_manifest_pair(const swb_manifest* m, int64_t k, const char** a, int64_t* alen, const char** b, int64_t* blen)
{
    if (!m || k < 0 || (size_t)k >= m->a.size()) return SWB_ERR_ARG;
    if (a) *a = m->a[(size_t)k].data();
    if (alen) *alen = (int64_t)m->a[(size_t)k].size();
    if (b) *b = m->b[(size_t)k].data();
    if (blen) *blen = (int64_t)m->b[(size_t)k].size();
    return SWB_OK;
}

void swb_manifest_free(swb_manifest* m) { delete m; }

}  // extern "C"
