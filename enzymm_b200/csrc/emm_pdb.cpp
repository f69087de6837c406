// emm_pdb.cpp -- native PDB ingest for the matching path (replaces what
// pyjess.Molecule.load does in C inside Jess, call site enzymm/jess_run.py:538).
//
// Fixed-column reader: ATOM and HETATM records, in file order, up to the first ENDMDL (SURVEY 8c
// rule 1); coordinates go through strtod, i.e. they are exactly the doubles Python's float()
// yields for the same text.  emm_pdb_load_files reads and parses many files on a thread pool into
// one SoA batch, which is what the batched upload wants.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/enzymm_b200.h"

namespace {

thread_local std::string t_error;

inline bool is_coord_record(const char *p, int64_t n)
{
    return n >= 6 && (memcmp(p, "ATOM  ", 6) == 0 || memcmp(p, "HETATM", 6) == 0);
}

// copy columns [a, b) of a line (clipped to its length) stripped of blanks, NUL padded to width
inline void field(const char *line, int64_t len, int a, int b, char *dst, int width)
{
    memset(dst, 0, (size_t)width);
    int lo = a, hi = std::min<int64_t>(b, len);
    while (lo < hi && (line[lo] == ' ' || line[lo] == '\t')) ++lo;
    while (hi > lo && (line[hi - 1] == ' ' || line[hi - 1] == '\t' || line[hi - 1] == '\r')) --hi;
    for (int i = 0; lo + i < hi && i < width; ++i) dst[i] = line[lo + i];
}

inline bool parse_int(const char *line, int64_t len, int a, int b, int32_t *out)
{
    char buf[16];
    field(line, len, a, b, buf, 15);
    buf[15] = 0;
    if (!buf[0]) return false;
    char *end = nullptr;
    const long v = strtol(buf, &end, 10);
    if (*end) return false;
    *out = (int32_t)v;
    return true;
}

// Plain fixed-point decimals ("-12.345") are converted as mantissa / 10^k with one correctly
// rounded division: mantissa < 2^53 and 10^k (k <= 18) are exact doubles, so the quotient is the
// double nearest to the decimal -- bit-identical to strtod.  Anything else falls back to strtod.
inline bool fast_real(const char *p, const char *end, double *out)
{
    static const double pow10[19] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12,
                                     1e13, 1e14, 1e15, 1e16, 1e17, 1e18};
    bool neg = false;
    if (p < end && (*p == '-' || *p == '+')) { neg = *p == '-'; ++p; }
    uint64_t mant = 0;
    int digits = 0, frac = 0;
    bool dot = false;
    for (; p < end; ++p) {
        if (*p >= '0' && *p <= '9') {
            mant = mant * 10 + (uint64_t)(*p - '0');
            ++digits;
            if (dot) ++frac;
        } else if (*p == '.' && !dot) {
            dot = true;
        } else {
            return false;
        }
    }
    if (digits == 0 || digits > 15 || frac > 18) return false;
    const double v = (double)mant / pow10[frac];
    *out = neg ? -v : v;
    return true;
}

inline bool parse_real(const char *line, int64_t len, int a, int b, double *out, bool optional)
{
    {
        int lo = a, hi = (int)std::min<int64_t>(b, len);
        while (lo < hi && line[lo] == ' ') ++lo;
        while (hi > lo && (line[hi - 1] == ' ' || line[hi - 1] == '\r')) --hi;
        if (lo == hi) { *out = 0.0; return optional; }
        if (fast_real(line + lo, line + hi, out)) return true;
    }
    char buf[24];
    field(line, len, a, b, buf, 23);
    buf[23] = 0;
    if (!buf[0]) { *out = 0.0; return optional; }
    char *end = nullptr;
    *out = strtod(buf, &end);
    return *end == 0;
}

inline bool fast_int(const char *p, const char *e, int32_t *out)
{
    while (p < e && *p == ' ') ++p;
    while (e > p && e[-1] == ' ') --e;
    bool neg = false;
    if (p < e && (*p == '-' || *p == '+')) { neg = *p == '-'; ++p; }
    if (p == e || e - p > 9) return false;
    int32_t v = 0;
    for (; p < e; ++p) {
        const unsigned d = (unsigned)(*p - '0');
        if (d > 9u) return false;
        v = v * 10 + (int32_t)d;
    }
    *out = neg ? -v : v;
    return true;
}

inline bool int_field(const char *line, int64_t ll, int a, int b, int32_t *out)
{
    return fast_int(line + a, line + std::min<int64_t>(b, ll), out) || parse_int(line, ll, a, b, out);
}

// The canonical "%W.Ff" layout (blanks, optional '-', digits, '.', F digits) without the generic
// scanner: same mantissa and the same single division as fast_real, so the same double.
template <int W, int F>
inline bool real_layout(const char *p, double *out)
{
    static_assert(F >= 1 && F <= 3 && W - F - 2 >= 0, "layout");
    if (p[W - F - 1] != '.') return false;
    int i = W - F - 2;
    unsigned d = (unsigned)(unsigned char)p[i] - '0';
    if (d > 9u) return false;
    uint64_t ip = d, scale = 10;
    bool neg = false;
    for (--i; i >= 0; --i) {
        const unsigned c = (unsigned char)p[i];
        d = c - '0';
        if (d <= 9u) { ip += d * scale; scale *= 10; }
        else if (c == '-') { neg = true; --i; break; }
        else if (c == ' ') break;
        else return false;
    }
    for (; i >= 0; --i)
        if (p[i] != ' ') return false;
    uint64_t frac = 0;
    for (int j = W - F; j < W; ++j) {
        d = (unsigned)(unsigned char)p[j] - '0';
        if (d > 9u) return false;
        frac = frac * 10 + d;
    }
    constexpr double div = F == 1 ? 1e1 : F == 2 ? 1e2 : 1e3;
    constexpr uint64_t mul = F == 1 ? 10 : F == 2 ? 100 : 1000;
    const double v = (double)(ip * mul + frac) / div;
    *out = neg ? -v : v;
    return true;
}

struct Columns {
    int32_t *serial; char *name; char *altloc; char *resname; char *chain; int32_t *resnum; char *icode;
    double *xyz; double *occupancy; double *bfactor; char *segment; char *element; int8_t *charge;
};

int64_t count_atoms(const char *text, int64_t len)
{
    int64_t n = 0, pos = 0;
    while (pos < len) {
        const char *nl = (const char *)memchr(text + pos, '\n', (size_t)(len - pos));
        const int64_t end = nl ? nl - text : len;
        const char *line = text + pos;
        const int64_t ll = end - pos;
        if (is_coord_record(line, ll)) ++n;
        else if (ll >= 6 && memcmp(line, "ENDMDL", 6) == 0) break;
        pos = end + 1;
    }
    return n;
}

// returns atoms parsed, or -1 (t_error set)
int64_t parse_into(const char *text, int64_t len, const Columns &c, int64_t base, int64_t capacity, char header_id[5])
{
    int64_t n = 0, pos = 0;
    bool have_header = false;
    memset(header_id, 0, 5);
    while (pos < len) {
        const char *nl = (const char *)memchr(text + pos, '\n', (size_t)(len - pos));
        const int64_t end = nl ? nl - text : len;
        const char *line = text + pos;
        int64_t ll = end - pos;
        while (ll > 0 && line[ll - 1] == '\r') --ll;
        if (is_coord_record(line, ll)) {
            if (n >= capacity) { t_error = "atom capacity exceeded"; return -1; }
            const int64_t i = base + n;
            double x, y, z;
            const bool wide = ll >= 66;
            if (ll < 54 || !int_field(line, ll, 6, 11, c.serial + i) || !int_field(line, ll, 22, 26, c.resnum + i) ||
                !(real_layout<8, 3>(line + 30, &x) || parse_real(line, ll, 30, 38, &x, false)) ||
                !(real_layout<8, 3>(line + 38, &y) || parse_real(line, ll, 38, 46, &y, false)) ||
                !(real_layout<8, 3>(line + 46, &z) || parse_real(line, ll, 46, 54, &z, false)) ||
                !((wide && real_layout<6, 2>(line + 54, c.occupancy + i)) || parse_real(line, ll, 54, 60, c.occupancy + i, true)) ||
                !((wide && real_layout<6, 2>(line + 60, c.bfactor + i)) || parse_real(line, ll, 60, 66, c.bfactor + i, true))) {
                t_error = "malformed PDB coordinate record: " + std::string(line, (size_t)std::min<int64_t>(ll, 80));
                return -1;
            }
            c.xyz[3 * i] = x; c.xyz[3 * i + 1] = y; c.xyz[3 * i + 2] = z;
            field(line, ll, 12, 16, c.name + 4 * i, 4);
            c.altloc[i] = ll > 16 ? line[16] : ' ';
            field(line, ll, 17, 20, c.resname + 4 * i, 4);
            field(line, ll, 20, 22, c.chain + 2 * i, 2);
            c.icode[i] = ll > 26 ? line[26] : ' ';
            field(line, ll, 72, 76, c.segment + 4 * i, 4);
            field(line, ll, 76, 78, c.element + 2 * i, 2);
            char chg[3];
            field(line, ll, 78, 80, chg, 2);
            chg[2] = 0;
            int8_t q = 0;
            if (chg[0] >= '0' && chg[0] <= '9') q = (int8_t)((chg[0] - '0') * ((chg[1] == '-') ? -1 : 1));
            c.charge[i] = q;
            ++n;
        } else if (ll >= 6 && memcmp(line, "ENDMDL", 6) == 0) {
            break;
        } else if (!have_header && ll >= 6 && memcmp(line, "HEADER", 6) == 0) {
            have_header = true;
            char id[5];
            field(line, ll, 62, 66, id, 4);
            id[4] = 0;
            memcpy(header_id, id, 5);
        }
        pos = end + 1;
    }
    return n;
}

// ---- files -> emm_batch columns -------------------------------------------------------------------

// blank-stripped field of width w (<= 4) as little-endian bytes, NUL padded: the same bytes field() yields
inline uint32_t strip_field(const char *p, int w)
{
    int lo = 0, hi = w;
    while (lo < hi && (p[lo] == ' ' || p[lo] == '\t')) ++lo;
    while (hi > lo && (p[hi - 1] == ' ' || p[hi - 1] == '\t' || p[hi - 1] == '\r')) --hi;
    uint32_t v = 0;
    for (int i = 0; lo + i < hi; ++i) v |= (uint32_t)(unsigned char)p[lo + i] << (8 * i);
    return v;
}

struct PackedCols {
    double *xyz; uint32_t *kind; int32_t *residue; float *bfactor; uint16_t *chain;
};

// open-addressing table of the distinct (resname, name) kinds of one file, first-appearance order
struct KindTable {
    std::vector<uint64_t> keys;          // in order of first appearance
    std::vector<uint32_t> slot;          // hash slots: index + 1, 0 = empty
    KindTable() : slot(1024, 0) {}
    void clear() { keys.clear(); std::fill(slot.begin(), slot.end(), 0u); }
    uint32_t lookup(uint64_t key)
    {
        const size_t mask = slot.size() - 1;
        size_t h = (size_t)((key * 0x9E3779B97F4A7C15ull) >> 40) & mask;
        for (;; h = (h + 1) & mask) {
            const uint32_t v = slot[h];
            if (!v) break;
            if (keys[v - 1] == key) return v - 1;
        }
        keys.push_back(key);
        slot[h] = (uint32_t)keys.size();
        if (keys.size() * 2 > slot.size()) {                  // grow and rehash
            std::vector<uint32_t> bigger(slot.size() * 4, 0u);
            const size_t m2 = bigger.size() - 1;
            for (size_t i = 0; i < keys.size(); ++i) {
                size_t g = (size_t)((keys[i] * 0x9E3779B97F4A7C15ull) >> 40) & m2;
                while (bigger[g]) g = (g + 1) & m2;
                bigger[g] = (uint32_t)i + 1;
            }
            slot.swap(bigger);
        }
        return (uint32_t)keys.size() - 1;
    }
};

// Residue names Match.query_residue_count counts (enzymm/utils.py:116-161: the 20 proteinogenic amino
// acids + EnzyMM's "special" ones), as blank-stripped little-endian packed names.
inline bool counted_residue(uint32_t resname)
{
    static const char *const names[] = {"ALA", "ARG", "ASN", "ASP", "CYS", "GLN", "GLU", "GLY", "HIS", "ILE", "LEU", "LYS",
                                        "MET", "PHE", "PRO", "SER", "THR", "TRP", "TYR", "VAL", "ASX", "GLX", "SEC", "PYL",
                                        "UNK", "MSE", "SEP", "TPO", "PTR", "HYP", "CME", "CSO", "CSD", "PCA", "MLY", "DAL",
                                        "DAR", "DSG", "ORN", "PTM"};
    static uint32_t packed[40];
    static const bool ready = [] {
        for (int i = 0; i < 40; ++i)
            packed[i] = (uint32_t)(unsigned char)names[i][0] | ((uint32_t)(unsigned char)names[i][1] << 8) |
                        ((uint32_t)(unsigned char)names[i][2] << 16);
        return true;
    }();
    (void)ready;
    for (int i = 0; i < 40; ++i)
        if (packed[i] == resname) return true;
    return false;
}

// distinct residue NUMBERS among atoms of counted residues, chain ignored (jess_run.py:487-496)
inline int32_t distinct_count(std::vector<int32_t> &v)
{
    std::sort(v.begin(), v.end());
    return (int32_t)(std::unique(v.begin(), v.end()) - v.begin());
}

// One file into the packed columns at [base, base+capacity).  kind[] receives file-local kind
// indices; *split is set when a residue key reappears after another residue (the caller then
// regroups the file).  Returns atoms parsed or -1 (t_error set).
int64_t pack_into(const char *text, int64_t len, const PackedCols &c, int64_t base, int64_t capacity,
                  KindTable &kinds, std::vector<uint64_t> &run_keys, bool *split, char header_id[5],
                  int32_t *residue_count)
{
    std::vector<int32_t> counted;
    int64_t n = 0, pos = 0;
    bool have_header = false;
    memset(header_id, 0, 5);
    uint64_t prev_kind = ~0ull, prev_res = ~0ull;
    uint32_t prev_kind_idx = 0;
    int32_t run = -1;
    run_keys.clear();
    *split = false;
    while (pos < len) {
        const char *nl = (const char *)memchr(text + pos, '\n', (size_t)(len - pos));
        const int64_t end = nl ? nl - text : len;
        const char *line = text + pos;
        int64_t ll = end - pos;
        while (ll > 0 && line[ll - 1] == '\r') --ll;
        if (is_coord_record(line, ll)) {
            if (n >= capacity) { t_error = "atom capacity exceeded"; return -1; }
            const int64_t i = base + n;
            double x, y, z, occ, bf;
            int32_t serial, resnum;
            const bool wide = ll >= 66;
            if (ll < 54 || !int_field(line, ll, 6, 11, &serial) || !int_field(line, ll, 22, 26, &resnum) ||
                !(real_layout<8, 3>(line + 30, &x) || parse_real(line, ll, 30, 38, &x, false)) ||
                !(real_layout<8, 3>(line + 38, &y) || parse_real(line, ll, 38, 46, &y, false)) ||
                !(real_layout<8, 3>(line + 46, &z) || parse_real(line, ll, 46, 54, &z, false)) ||
                !((wide && real_layout<6, 2>(line + 54, &occ)) || parse_real(line, ll, 54, 60, &occ, true)) ||
                !((wide && real_layout<6, 2>(line + 60, &bf)) || parse_real(line, ll, 60, 66, &bf, true))) {
                t_error = "malformed PDB coordinate record: " + std::string(line, (size_t)std::min<int64_t>(ll, 80));
                return -1;
            }
            c.xyz[3 * i] = x; c.xyz[3 * i + 1] = y; c.xyz[3 * i + 2] = z;
            c.bfactor[i] = (float)bf;
            const uint32_t name = strip_field(line + 12, 4), resname = strip_field(line + 17, 3);
            const uint16_t chain = (uint16_t)strip_field(line + 20, 2);
            const uint64_t kkey = (uint64_t)resname | ((uint64_t)name << 32);
            const uint64_t rkey = ((uint64_t)chain << 32) | (uint32_t)resnum;
            // once per run of equal (residue name, chain, number), not once per atom
            if (((uint32_t)prev_kind != resname || rkey != prev_res) && counted_residue(resname)) counted.push_back(resnum);
            if (kkey != prev_kind) { prev_kind = kkey; prev_kind_idx = kinds.lookup(kkey); }
            c.kind[i] = prev_kind_idx;
            c.chain[i] = chain;
            if (rkey != prev_res || run < 0) {
                prev_res = rkey;
                ++run;
                run_keys.push_back(rkey);
            }
            c.residue[i] = run;
            ++n;
        } else if (ll >= 6 && memcmp(line, "ENDMDL", 6) == 0) {
            break;
        } else if (!have_header && ll >= 6 && memcmp(line, "HEADER", 6) == 0) {
            have_header = true;
            char id[5];
            field(line, ll, 62, 66, id, 4);
            id[4] = 0;
            memcpy(header_id, id, 5);
        }
        pos = end + 1;
    }
    {
        std::vector<uint64_t> sorted(run_keys);
        std::sort(sorted.begin(), sorted.end());
        *split = std::adjacent_find(sorted.begin(), sorted.end()) != sorted.end();
    }
    *residue_count = distinct_count(counted);
    return n;
}

// A file with a split residue: number residues by first appearance of their key and make every
// residue's atoms contiguous with a stable sort (packing.py residue_ordinals does the same).
void regroup_file(const PackedCols &c, int64_t base, int64_t n, std::vector<uint64_t> &run_keys, int32_t *atom_id)
{
    std::unordered_map<uint64_t, int32_t> rank_of;
    std::vector<int32_t> rank_of_run(run_keys.size());
    std::vector<uint64_t> by_rank;                 // residue key of every ordinal, first appearance order
    for (size_t r = 0; r < run_keys.size(); ++r) {
        const auto ins = rank_of.emplace(run_keys[r], (int32_t)rank_of.size());
        if (ins.second) by_rank.push_back(run_keys[r]);
        rank_of_run[r] = ins.first->second;
    }
    std::vector<int32_t> ordinal((size_t)n), order((size_t)n);
    for (int64_t j = 0; j < n; ++j) {
        ordinal[(size_t)j] = rank_of_run[(size_t)c.residue[base + j]];
        order[(size_t)j] = (int32_t)j;
    }
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return ordinal[(size_t)a] < ordinal[(size_t)b]; });
    std::vector<double> xyz((size_t)(3 * n));
    std::vector<uint32_t> kind((size_t)n);
    std::vector<float> bf((size_t)n);
    std::vector<uint16_t> chain((size_t)n);
    for (int64_t j = 0; j < n; ++j) {
        const int64_t src = base + order[(size_t)j];
        xyz[(size_t)(3 * j)] = c.xyz[3 * src]; xyz[(size_t)(3 * j + 1)] = c.xyz[3 * src + 1]; xyz[(size_t)(3 * j + 2)] = c.xyz[3 * src + 2];
        kind[(size_t)j] = c.kind[src]; bf[(size_t)j] = c.bfactor[src]; chain[(size_t)j] = c.chain[src];
    }
    for (int64_t j = 0; j < n; ++j) {
        const int64_t dst = base + j;
        c.xyz[3 * dst] = xyz[(size_t)(3 * j)]; c.xyz[3 * dst + 1] = xyz[(size_t)(3 * j + 1)]; c.xyz[3 * dst + 2] = xyz[(size_t)(3 * j + 2)];
        c.kind[dst] = kind[(size_t)j]; c.bfactor[dst] = bf[(size_t)j]; c.chain[dst] = chain[(size_t)j];
        c.residue[dst] = ordinal[(size_t)order[(size_t)j]];
        atom_id[dst] = order[(size_t)j];
    }
    run_keys.swap(by_rank);
}

bool read_file(const char *path, std::string &out, int *err)
{
    FILE *f = fopen(path, "rb");
    if (!f) { *err = 1; return false; }
    if (fseek(f, 0, SEEK_END) != 0) { fclose(f); *err = 2; return false; }   // directories fail here or at read
    const long size = ftell(f);
    if (size < 0) { fclose(f); *err = 2; return false; }
    rewind(f);
    out.resize((size_t)size);
    const size_t got = size ? fread(&out[0], 1, (size_t)size, f) : 0;
    fclose(f);
    if (got != (size_t)size) { *err = 2; return false; }
    return true;
}

// whole file into a reusable buffer (grown, never shrunk: no mmap churn between files)
bool read_file_into(const char *path, std::vector<char> &buf, int64_t *len, int *err)
{
    const int fd = open(path, O_RDONLY | O_CLOEXEC);
    if (fd < 0) { *err = 1; return false; }
    struct stat st;
    if (fstat(fd, &st) != 0 || S_ISDIR(st.st_mode)) { close(fd); *err = 2; return false; }
    size_t size = (size_t)st.st_size, got = 0;
    if (buf.size() < size + 1) buf.resize(size + size / 4 + 4096);
    for (;;) {
        if (got == buf.size()) buf.resize(buf.size() * 2);        // the file grew, or st_size lied (procfs)
        const ssize_t r = read(fd, buf.data() + got, buf.size() - got);
        if (r < 0) { close(fd); *err = 2; return false; }
        if (r == 0) break;
        got += (size_t)r;
    }
    close(fd);
    *len = (int64_t)got;
    return true;
}

// the packed columns of one file, in one allocation
struct FileBlock {
    int64_t n = 0;
    std::unique_ptr<char[]> mem;
    std::unique_ptr<int32_t[]> atom_id;
    double *xyz = nullptr; uint32_t *kind = nullptr; int32_t *residue = nullptr; float *bfactor = nullptr;
    uint16_t *chain = nullptr;
    std::vector<uint64_t> kinds;
    std::vector<uint64_t> res_keys;      // (chain << 32 | residue number) of every residue ordinal
    int32_t residue_count = 0;           // Match.query_residue_count of this structure
    bool split = false;
    void allocate(int64_t count)
    {
        n = count;
        const size_t c = (size_t)std::max<int64_t>(count, 1);
        mem.reset(new char[c * (24 + 4 + 4 + 4 + 2) + 64]);
        char *p = mem.get();
        p += (8 - (reinterpret_cast<uintptr_t>(p) & 7)) & 7;
        xyz = reinterpret_cast<double *>(p); p += c * 24;
        kind = reinterpret_cast<uint32_t *>(p); p += c * 4;
        residue = reinterpret_cast<int32_t *>(p); p += c * 4;
        bfactor = reinterpret_cast<float *>(p); p += c * 4;
        chain = reinterpret_cast<uint16_t *>(p);
    }
    void release() { mem.reset(); atom_id.reset(); std::vector<uint64_t>().swap(kinds); std::vector<uint64_t>().swap(res_keys); }
};

}  // namespace

// array that is NOT value-initialised: its pages are first touched by the worker that fills them
template <typename T>
struct RawArray {
    std::unique_ptr<T[]> p;
    size_t n = 0;
    void resize(size_t count) { p.reset(count ? new T[count] : nullptr); n = count; }
    T *data() { return p.get(); }
    const T *data() const { return p.get(); }
    T &operator[](size_t i) { return p[i]; }
    const T &operator[](size_t i) const { return p[i]; }
};

struct emm_pdb_batch {
    int32_t n_files = 0;
    int64_t n_atoms = 0;
    std::vector<int64_t> atom_off;
    // Molecule columns (emm_pdb_load_files): not value-initialised, every entry is written by the parser
    RawArray<int32_t> serial, resnum;
    RawArray<char> name, altloc, resname, chain, icode, segment, element;
    std::vector<char> header_id;
    RawArray<double> xyz, occupancy, bfactor;
    RawArray<int8_t> charge;
    // packed form (emm_pdb_pack_files)
    bool packed = false, has_atom_id = false, has_klass = false;
    int n_threads = 1;
    RawArray<uint16_t> klass;
    RawArray<double> pxyz;
    RawArray<uint32_t> kind;
    RawArray<int32_t> residue, atom_id;
    RawArray<float> bfactor32;
    RawArray<uint16_t> chain16;
    std::vector<char> kind_names;
    std::vector<int64_t> res_off;        // [n_files+1] into res_key
    std::vector<uint64_t> res_key;       // per residue ordinal: chain << 32 | residue number
    std::vector<int32_t> residue_count;  // per file
};

extern "C" {

const char *emm_pdb_last_error(void) { return t_error.c_str(); }

int emm_pdb_count_atoms(const char *text, int64_t len, int64_t *n_atoms)
{
    if (!text || !n_atoms || len < 0) return EMM_ERR_INVALID;
    *n_atoms = count_atoms(text, len);
    return EMM_OK;
}

int emm_pdb_parse(const char *text, int64_t len, int64_t capacity, int32_t *serial, char *name, char *altloc,
                  char *resname, char *chain, int32_t *resnum, char *icode, double *xyz, double *occupancy,
                  double *bfactor, char *segment, char *element, int8_t *charge, char *header_id, int64_t *n_atoms)
{
    if (!text || !n_atoms || len < 0 || !header_id) return EMM_ERR_INVALID;
    Columns c{serial, name, altloc, resname, chain, resnum, icode, xyz, occupancy, bfactor, segment, element, charge};
    const int64_t n = parse_into(text, len, c, 0, capacity, header_id);
    if (n < 0) return EMM_ERR_INPUT;
    *n_atoms = n;
    return EMM_OK;
}

int emm_pdb_load_files(const char *const *paths, int32_t n_files, int32_t n_threads, emm_pdb_batch **out)
{
    if (!paths || !out || n_files < 0) return EMM_ERR_INVALID;
    *out = nullptr;
    emm_pdb_batch *b = new emm_pdb_batch();
    b->n_files = n_files;
    std::vector<std::string> texts((size_t)n_files);
    std::vector<int64_t> counts((size_t)n_files, 0);
    std::vector<int> errs((size_t)n_files, 0);
    if (n_threads < 1) n_threads = 1;
    n_threads = std::min<int32_t>(n_threads, std::max(n_files, 1));
    {
        std::atomic<int> next(0);
        auto work = [&]() {
            for (int i; (i = next.fetch_add(1)) < n_files;) {
                if (read_file(paths[i], texts[(size_t)i], &errs[(size_t)i]))
                    counts[(size_t)i] = count_atoms(texts[(size_t)i].data(), (int64_t)texts[(size_t)i].size());
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
        work();
        for (auto &t : pool) t.join();
    }
    for (int i = 0; i < n_files; ++i)
        if (errs[(size_t)i]) {
            t_error = std::string(errs[(size_t)i] == 1 ? "cannot open " : "cannot read ") + paths[i];
            const int rc = errs[(size_t)i] == 1 ? EMM_ERR_INVALID : EMM_ERR_INPUT;
            delete b;
            return rc;
        }
    b->atom_off.assign((size_t)n_files + 1, 0);
    for (int i = 0; i < n_files; ++i) b->atom_off[(size_t)i + 1] = b->atom_off[(size_t)i] + counts[(size_t)i];
    const size_t n = (size_t)b->atom_off[(size_t)n_files];
    b->n_atoms = (int64_t)n;
    b->serial.resize(n); b->resnum.resize(n); b->name.resize(4 * n); b->altloc.resize(n); b->resname.resize(4 * n);
    b->chain.resize(2 * n); b->icode.resize(n); b->segment.resize(4 * n); b->element.resize(2 * n);
    b->xyz.resize(3 * n); b->occupancy.resize(n); b->bfactor.resize(n); b->charge.resize(n);
    b->header_id.assign(5 * (size_t)n_files, 0);
    Columns c{b->serial.data(), b->name.data(), b->altloc.data(), b->resname.data(), b->chain.data(), b->resnum.data(),
              b->icode.data(), b->xyz.data(), b->occupancy.data(), b->bfactor.data(), b->segment.data(),
              b->element.data(), b->charge.data()};
    std::atomic<int> next(0), failed(-1);
    std::vector<std::string> messages((size_t)n_files);
    auto work = [&]() {
        for (int i; (i = next.fetch_add(1)) < n_files;) {
            const int64_t got = parse_into(texts[(size_t)i].data(), (int64_t)texts[(size_t)i].size(), c,
                                           b->atom_off[(size_t)i], counts[(size_t)i], &b->header_id[5 * (size_t)i]);
            if (got != counts[(size_t)i]) { messages[(size_t)i] = t_error; failed.store(i); }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
    if (failed.load() >= 0) {
        t_error = std::string(paths[failed.load()]) + ": " + messages[(size_t)failed.load()];
        delete b;
        return EMM_ERR_INPUT;
    }
    *out = b;
    return EMM_OK;
}

int emm_pdb_batch_columns(const emm_pdb_batch *b, emm_pdb_columns *out)
{
    if (!b || !out || b->packed) return EMM_ERR_INVALID;
    out->n_files = b->n_files;
    out->n_atoms = b->n_atoms;
    out->atom_off = b->atom_off.data();
    out->serial = b->serial.data(); out->name = b->name.data(); out->altloc = b->altloc.data();
    out->resname = b->resname.data(); out->chain = b->chain.data(); out->resnum = b->resnum.data();
    out->icode = b->icode.data(); out->xyz = b->xyz.data(); out->occupancy = b->occupancy.data();
    out->bfactor = b->bfactor.data(); out->segment = b->segment.data(); out->element = b->element.data();
    out->charge = b->charge.data(); out->header_id = b->header_id.data();
    return EMM_OK;
}

// Blocks (one per structure, already parsed) -> the batch columns: offsets, the merged kind table
// (structure order: deterministic whatever the thread count) and a parallel copy into place.
static void finish_packed(emm_pdb_batch *b, std::vector<FileBlock> &blocks, int n_threads)
{
    const size_t nf = blocks.size();
    const int n_files = (int)nf;
    auto run_pool = [&](auto &&work) {
        std::vector<std::thread> pool;
        for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
        work();
        for (auto &t : pool) t.join();
    };
    b->atom_off.assign(nf + 1, 0);
    bool any_split = false;
    for (size_t i = 0; i < nf; ++i) {
        b->atom_off[i + 1] = b->atom_off[i] + blocks[i].n;
        any_split = any_split || blocks[i].split;
    }
    const size_t n = (size_t)b->atom_off[nf];
    b->n_atoms = (int64_t)n;
    b->pxyz.resize(3 * n); b->kind.resize(n); b->residue.resize(n); b->bfactor32.resize(n); b->chain16.resize(n);
    b->has_atom_id = any_split;
    if (any_split) b->atom_id.resize(n);
    // merge the per-file kind lists in file order (deterministic whatever the thread count)
    KindTable global;
    std::vector<std::vector<uint32_t>> remap(nf);
    for (size_t f = 0; f < nf; ++f) {
        remap[f].resize(blocks[f].kinds.size());
        for (size_t j = 0; j < blocks[f].kinds.size(); ++j) remap[f][j] = global.lookup(blocks[f].kinds[j]);
    }
    b->kind_names.assign(8 * global.keys.size(), 0);
    for (size_t j = 0; j < global.keys.size(); ++j) memcpy(&b->kind_names[8 * j], &global.keys[j], 8);
    b->res_off.assign(nf + 1, 0);
    b->residue_count.assign(nf, 0);
    for (size_t f = 0; f < nf; ++f) {
        b->res_off[f + 1] = b->res_off[f] + (int64_t)blocks[f].res_keys.size();
        b->residue_count[f] = blocks[f].residue_count;
    }
    b->res_key.resize((size_t)b->res_off[nf]);
    for (size_t f = 0; f < nf; ++f)
        std::copy(blocks[f].res_keys.begin(), blocks[f].res_keys.end(), b->res_key.begin() + b->res_off[f]);
    // pass 2: blocks -> their place in the batch columns (first touch of those pages, in parallel)
    {
        std::atomic<int> next(0);
        emm_pdb_batch *bp = b;
        run_pool([&]() {
            for (int i; (i = next.fetch_add(1)) < n_files;) {
                const size_t f = (size_t)i;
                FileBlock &blk = blocks[f];
                const size_t lo = (size_t)bp->atom_off[f], cnt = (size_t)blk.n;
                if (cnt) {
                    memcpy(bp->pxyz.data() + 3 * lo, blk.xyz, cnt * 3 * sizeof(double));
                    memcpy(bp->residue.data() + lo, blk.residue, cnt * sizeof(int32_t));
                    memcpy(bp->bfactor32.data() + lo, blk.bfactor, cnt * sizeof(float));
                    memcpy(bp->chain16.data() + lo, blk.chain, cnt * sizeof(uint16_t));
                    const std::vector<uint32_t> &m = remap[f];
                    uint32_t *kd = bp->kind.data() + lo;
                    for (size_t a = 0; a < cnt; ++a) kd[a] = m[blk.kind[a]];
                    if (bp->has_atom_id) {
                        int32_t *ad = bp->atom_id.data() + lo;
                        if (blk.split) memcpy(ad, blk.atom_id.get(), cnt * sizeof(int32_t));
                        else for (size_t a = 0; a < cnt; ++a) ad[a] = (int32_t)a;
                    }
                }
                blk.release();
            }
        });
    }
}

int emm_pdb_pack_files(const char *const *paths, int32_t n_files, int32_t n_threads, emm_pdb_batch **out)
{
    if (!paths || !out || n_files < 0) return EMM_ERR_INVALID;
    *out = nullptr;
    std::unique_ptr<emm_pdb_batch> b(new emm_pdb_batch());
    b->n_files = n_files;
    b->packed = true;
    const size_t nf = (size_t)n_files;
    if (n_threads < 1) n_threads = 1;
    n_threads = std::min<int32_t>(n_threads, std::max(n_files, 1));
    b->n_threads = n_threads;
    auto run_pool = [&](auto &&work) {
        std::vector<std::thread> pool;
        for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
        work();
        for (auto &t : pool) t.join();
    };
    // pass 1, one file at a time per worker: read into the worker's reusable text buffer, count,
    // parse into a block sized for this file while the text is still cache-hot
    std::vector<FileBlock> blocks(nf);
    std::vector<int> errs(nf, 0);
    std::vector<std::string> messages(nf);
    b->header_id.assign(5 * nf, 0);
    {
        std::atomic<int> next(0);
        run_pool([&]() {
            KindTable kinds;
            std::vector<uint64_t> run_keys;
            std::vector<char> text;
            for (int i; (i = next.fetch_add(1)) < n_files;) {
                const size_t f = (size_t)i;
                int64_t len = 0;
                if (!read_file_into(paths[i], text, &len, &errs[f])) continue;
                FileBlock &blk = blocks[f];
                const int64_t count = count_atoms(text.data(), len);
                blk.allocate(count);
                kinds.clear();
                const PackedCols c{blk.xyz, blk.kind, blk.residue, blk.bfactor, blk.chain};
                const int64_t got = pack_into(text.data(), len, c, 0, count, kinds, run_keys, &blk.split, &b->header_id[5 * f],
                                              &blk.residue_count);
                if (got != count) { messages[f] = t_error; errs[f] = 3; continue; }
                blk.kinds = kinds.keys;
                if (blk.split) {
                    blk.atom_id.reset(new int32_t[(size_t)std::max<int64_t>(count, 1)]);
                    regroup_file(c, 0, count, run_keys, blk.atom_id.get());
                }
                blk.res_keys = run_keys;
            }
        });
    }
    for (int i = 0; i < n_files; ++i)
        if (errs[(size_t)i]) {
            const int e = errs[(size_t)i];
            if (e == 3) t_error = std::string(paths[i]) + ": " + messages[(size_t)i];
            else t_error = std::string(e == 1 ? "cannot open " : "cannot read ") + paths[i];
            return e == 1 ? EMM_ERR_INVALID : EMM_ERR_INPUT;
        }
    finish_packed(b.get(), blocks, n_threads);
    *out = b.release();
    return EMM_OK;
}

int emm_pack_columns(int32_t n_structures, const int64_t *sizes, const uint8_t *const *name4,
                     const uint8_t *const *resname4, const uint8_t *const *chain2, const int32_t *const *resnum,
                     const double *const *xyz, const double *const *bfactor, int32_t n_threads, emm_pdb_batch **out)
{
    if (!out || n_structures < 0 || (n_structures > 0 && (!sizes || !name4 || !resname4 || !chain2 || !resnum || !xyz || !bfactor)))
        return EMM_ERR_INVALID;
    *out = nullptr;
    std::unique_ptr<emm_pdb_batch> b(new emm_pdb_batch());
    b->n_files = n_structures;
    b->packed = true;
    const size_t nf = (size_t)n_structures;
    if (n_threads < 1) n_threads = 1;
    n_threads = std::min<int32_t>(n_threads, std::max(n_structures, 1));
    b->n_threads = n_threads;
    b->header_id.assign(5 * nf, 0);
    std::vector<FileBlock> blocks(nf);
    {
        std::atomic<int> next(0);
        auto work = [&]() {
            KindTable kinds;
            std::vector<uint64_t> run_keys;
            for (int i; (i = next.fetch_add(1)) < n_structures;) {
                const size_t f = (size_t)i;
                const int64_t n = sizes[f];
                FileBlock &blk = blocks[f];
                blk.allocate(n);
                kinds.clear();
                run_keys.clear();
                uint64_t prev_kind = ~0ull, prev_res = ~0ull;
                uint32_t prev_kind_idx = 0;
                int32_t run = -1;
                const uint8_t *nm = name4[f], *rn = resname4[f], *ch = chain2[f];
                std::vector<int32_t> counted;
                for (int64_t a = 0; a < n; ++a) {
                    uint32_t name, res;
                    uint16_t chain;
                    memcpy(&name, nm + 4 * a, 4);
                    memcpy(&res, rn + 4 * a, 4);
                    memcpy(&chain, ch + 2 * a, 2);
                    const uint64_t kkey = (uint64_t)res | ((uint64_t)name << 32);
                    const uint64_t rkey = ((uint64_t)chain << 32) | (uint32_t)resnum[f][a];
                    if (((uint32_t)prev_kind != res || rkey != prev_res) && counted_residue(res)) counted.push_back(resnum[f][a]);
                    if (kkey != prev_kind) { prev_kind = kkey; prev_kind_idx = kinds.lookup(kkey); }
                    blk.kind[a] = prev_kind_idx;
                    blk.chain[a] = chain;
                    if (rkey != prev_res || run < 0) { prev_res = rkey; ++run; run_keys.push_back(rkey); }
                    blk.residue[a] = run;
                    blk.bfactor[a] = (float)bfactor[f][a];
                }
                if (n) memcpy(blk.xyz, xyz[f], (size_t)n * 3 * sizeof(double));
                blk.kinds = kinds.keys;
                std::vector<uint64_t> sorted(run_keys);
                std::sort(sorted.begin(), sorted.end());
                blk.split = std::adjacent_find(sorted.begin(), sorted.end()) != sorted.end();
                if (blk.split) {
                    blk.atom_id.reset(new int32_t[(size_t)std::max<int64_t>(n, 1)]);
                    const PackedCols c{blk.xyz, blk.kind, blk.residue, blk.bfactor, blk.chain};
                    regroup_file(c, 0, n, run_keys, blk.atom_id.get());
                }
                blk.res_keys = run_keys;
                blk.residue_count = distinct_count(counted);
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
        work();
        for (auto &t : pool) t.join();
    }
    finish_packed(b.get(), blocks, n_threads);
    *out = b.release();
    return EMM_OK;
}

int emm_pdb_batch_classify(emm_pdb_batch *b, const uint16_t *class_of_kind, int32_t n_kinds)
{
    if (!b || !b->packed || !class_of_kind || n_kinds < (int32_t)(b->kind_names.size() / 8)) return EMM_ERR_INVALID;
    const size_t n = (size_t)b->n_atoms;
    b->klass.resize(n);
    const int n_threads = std::max(1, std::min<int>(b->n_threads, (int)(n / 65536) + 1));
    const size_t per = (n + (size_t)n_threads - 1) / (size_t)n_threads;
    auto work = [&](int t) {
        const size_t lo = std::min(n, per * (size_t)t), hi = std::min(n, lo + per);
        const uint32_t *kd = b->kind.data();
        uint16_t *out = b->klass.data();
        for (size_t a = lo; a < hi; ++a) out[a] = class_of_kind[kd[a]];
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto &t : pool) t.join();
    b->has_klass = true;
    return EMM_OK;
}

int emm_pdb_batch_packed(const emm_pdb_batch *b, emm_pdb_packed *out)
{
    if (!b || !out || !b->packed) return EMM_ERR_INVALID;
    out->n_files = b->n_files;
    out->n_atoms = b->n_atoms;
    out->atom_off = b->atom_off.data();
    out->xyz = b->pxyz.data(); out->kind = b->kind.data(); out->residue = b->residue.data();
    out->bfactor = b->bfactor32.data(); out->chain = b->chain16.data();
    out->atom_id = b->has_atom_id ? b->atom_id.data() : nullptr;
    out->klass = b->has_klass ? b->klass.data() : nullptr;
    out->n_kinds = (int32_t)(b->kind_names.size() / 8);
    out->kind_names = b->kind_names.data();
    out->header_id = b->header_id.data();
    out->res_off = b->res_off.data();
    out->res_key = b->res_key.data();
    out->residue_count = b->residue_count.data();
    return EMM_OK;
}

void emm_pdb_batch_free(emm_pdb_batch *b) { delete b; }

}  // extern "C"
