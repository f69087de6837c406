// emm_pdb.cpp -- native PDB ingest for the matching path (replaces what
// pyjess.Molecule.load does in C inside Jess, call site enzymm/jess_run.py:538).
//
// Fixed-column reader: ATOM and HETATM records, in file order, up to the first ENDMDL (SURVEY 8c
// rule 1); coordinates go through strtod, i.e. they are exactly the doubles Python's float()
// yields for the same text.  emm_pdb_load_files reads and parses many files on a thread pool into
// one SoA batch, which is what the batched upload wants.
//
// mmCIF: a file whose first token is a data_ block header is read through its _atom_site category
// instead (pyjess.Molecule.load(format="detect") accepts both): rows of the first model only, the
// label_* identifiers unless EMM_PDB_CIF_AUTHOR asks for auth_*, numbers converted exactly as in the
// PDB path.  No reference test holds a CIF file, so which identifiers PyJess would pick is unpinned.
// gzip-compressed files (magic 1f 8b) are inflated on the worker that reads them.
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

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

// hot helpers of the fixed-column reader: inlined into its loops whatever else calls them
#define EMM_HOT inline __attribute__((always_inline))

thread_local std::string t_error;

EMM_HOT bool is_coord_record(const char *p, int64_t n)
{
    return n >= 6 && (memcmp(p, "ATOM  ", 6) == 0 || memcmp(p, "HETATM", 6) == 0);
}

// copy columns [a, b) of a line (clipped to its length) stripped of blanks, NUL padded to width
EMM_HOT void field(const char *line, int64_t len, int a, int b, char *dst, int width)
{
    memset(dst, 0, (size_t)width);
    int lo = a, hi = std::min<int64_t>(b, len);
    while (lo < hi && (line[lo] == ' ' || line[lo] == '\t')) ++lo;
    while (hi > lo && (line[hi - 1] == ' ' || line[hi - 1] == '\t' || line[hi - 1] == '\r')) --hi;
    for (int i = 0; lo + i < hi && i < width; ++i) dst[i] = line[lo + i];
}

EMM_HOT bool parse_int(const char *line, int64_t len, int a, int b, int32_t *out)
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
EMM_HOT bool fast_real(const char *p, const char *end, double *out)
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

EMM_HOT bool parse_real(const char *line, int64_t len, int a, int b, double *out, bool optional)
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

EMM_HOT bool fast_int(const char *p, const char *e, int32_t *out)
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

EMM_HOT bool int_field(const char *line, int64_t ll, int a, int b, int32_t *out)
{
    return fast_int(line + a, line + std::min<int64_t>(b, ll), out) || parse_int(line, ll, a, b, out);
}

// The canonical "%W.Ff" layout (blanks, optional '-', digits, '.', F digits) without the generic
// scanner: same mantissa and the same single division as fast_real, so the same double.
template <int W, int F>
EMM_HOT bool real_layout(const char *p, double *out)
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

int64_t pdb_count_atoms(const char *text, int64_t len)
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
int64_t pdb_parse_into(const char *text, int64_t len, const Columns &c, int64_t base, int64_t capacity, char header_id[5])
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
EMM_HOT uint32_t strip_field(const char *p, int w)
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

// The per-atom bookkeeping of the packed form, shared by the PDB reader, the mmCIF reader and
// emm_pack_columns: file-local kind index, chain code, residue runs (a run = consecutive atoms with
// one (chain, number) key), the residue numbers Match.query_residue_count counts.
// (Everything the object owns is a scalar and every method is inlined, so that it lives in registers
// inside the readers' loops; the vectors belong to the caller.)
struct PackState {
    const PackedCols c;
    KindTable &kinds;
    std::vector<uint64_t> &run_keys;
    std::vector<int32_t> &counted;
    uint64_t prev_kind = ~0ull, prev_res = ~0ull;
    uint32_t prev_kind_idx = 0;
    int32_t run = -1;
    PackState(const PackedCols &cols, KindTable &k, std::vector<uint64_t> &r, std::vector<int32_t> &n)
        : c(cols), kinds(k), run_keys(r), counted(n) { run_keys.clear(); counted.clear(); }
    // name / resname: blank-stripped little-endian packed bytes (strip_field)
    EMM_HOT void add(int64_t i, uint32_t name, uint32_t resname, uint16_t chain, int32_t resnum, float bf)
    {
        c.bfactor[i] = bf;
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
    }
    // *split: a residue key reappears after another residue (the caller then regroups the file)
    EMM_HOT void finish(bool *split, int32_t *residue_count)
    {
        std::vector<uint64_t> sorted(run_keys);
        std::sort(sorted.begin(), sorted.end());
        *split = std::adjacent_find(sorted.begin(), sorted.end()) != sorted.end();
        *residue_count = distinct_count(counted);
    }
};

// One file into the packed columns at [base, base+capacity).  kind[] receives file-local kind
// indices; *split is set when a residue key reappears after another residue (the caller then
// regroups the file).  Returns atoms parsed or -1 (t_error set).
int64_t pdb_pack_into(const char *text, int64_t len, const PackedCols &c, int64_t base, int64_t capacity,
                      KindTable &kinds, std::vector<uint64_t> &run_keys, bool *split, char header_id[5],
                      int32_t *residue_count)
{
    std::vector<int32_t> counted;
    PackState state(c, kinds, run_keys, counted);
    int64_t n = 0, pos = 0;
    bool have_header = false;
    memset(header_id, 0, 5);
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
            state.add(i, strip_field(line + 12, 4), strip_field(line + 17, 3), (uint16_t)strip_field(line + 20, 2), resnum, (float)bf);
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
    state.finish(split, residue_count);
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

// ---- mmCIF ----------------------------------------------------------------------------------------

struct CifSpaceTable {
    bool is[256];
    constexpr CifSpaceTable() : is()
    {
        for (int i = 0; i < 256; ++i) is[i] = i == ' ' || i == '\t' || i == '\r' || i == '\n';
    }
};
constexpr CifSpaceTable kCifSpace;
EMM_HOT bool cif_space(char ch) { return kCifSpace.is[(unsigned char)ch]; }

// CIF 1.1 tokens: bare words, '...' / "..." (the closing quote is the one followed by white space),
// ;-delimited text fields, # comments.
struct CifTok {
    const char *p;
    int64_t len;
    bool quoted;
    bool is(const char *word) const
    {
        const size_t n = strlen(word);
        return !quoted && (size_t)len == n && strncasecmp(p, word, n) == 0;
    }
    bool starts(const char *prefix) const
    {
        const size_t n = strlen(prefix);
        return !quoted && (size_t)len >= n && strncasecmp(p, prefix, n) == 0;
    }
    bool null() const { return !quoted && len == 1 && (p[0] == '.' || p[0] == '?'); }
};

struct CifScanner {
    const char *begin, *p, *end;
    CifScanner(const char *text, int64_t len) : begin(text), p(text), end(text + len) {}
    EMM_HOT bool next(CifTok &t)
    {
        for (;;) {
            while (p < end && cif_space(*p)) ++p;
            if (p >= end) return false;
            if (*p != '#') break;
            const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
            p = nl ? nl + 1 : end;
        }
        if (*p == ';' && (p == begin || p[-1] == '\n')) {
            const char *s = p + 1, *q = s;
            for (;;) {
                const char *nl = (const char *)memchr(q, '\n', (size_t)(end - q));
                if (!nl) { q = end; break; }
                if (nl + 1 < end && nl[1] == ';') { q = nl; break; }
                q = nl + 1;
            }
            t.p = s; t.len = q - s; t.quoted = true;
            p = q < end ? q + 2 : end;
            return true;
        }
        if (*p == '\'' || *p == '"') {
            const char quote = *p;
            const char *s = p + 1, *q = s;
            while (q < end && *q != '\n' && !(*q == quote && (q + 1 == end || cif_space(q[1])))) ++q;
            t.p = s; t.len = q - s; t.quoted = true;
            p = (q < end && *q == quote) ? q + 1 : q;
            return true;
        }
        const char *s = p;
        while (p < end && !cif_space(*p)) ++p;
        t.p = s; t.len = p - s; t.quoted = false;
        return true;
    }
};

// The first token of the text (after white space and comments) is a data_ block header.
bool looks_like_cif(const char *text, int64_t len)
{
    CifScanner scan(text, std::min<int64_t>(len, 65536));
    CifTok t;
    return scan.next(t) && t.starts("data_");
}

enum CifCol { C_GROUP, C_ID, C_SYMBOL, C_L_ATOM, C_A_ATOM, C_ALT, C_L_COMP, C_A_COMP, C_L_ASYM, C_A_ASYM, C_L_SEQ, C_A_SEQ,
              C_INS, C_X, C_Y, C_Z, C_OCC, C_B, C_CHARGE, C_MODEL, C_N };
const char *const kCifTags[C_N] = {"group_PDB", "id", "type_symbol", "label_atom_id", "auth_atom_id", "label_alt_id",
                                   "label_comp_id", "auth_comp_id", "label_asym_id", "auth_asym_id", "label_seq_id",
                                   "auth_seq_id", "pdbx_PDB_ins_code", "Cartn_x", "Cartn_y", "Cartn_z", "occupancy",
                                   "B_iso_or_equiv", "pdbx_formal_charge", "pdbx_PDB_model_num"};

inline int cif_column(const CifTok &tag)          // "_atom_site.<item>" -> CifCol or -1
{
    const int64_t skip = 11;
    for (int k = 0; k < C_N; ++k) {
        const size_t n = strlen(kCifTags[k]);
        if ((size_t)(tag.len - skip) == n && strncasecmp(tag.p + skip, kCifTags[k], n) == 0) return k;
    }
    return -1;
}

struct AtomRec {
    int32_t serial, resnum;
    uint32_t name, resname;          // blank-stripped little-endian packed bytes, as strip_field
    uint16_t chain, element;
    char altloc, icode;
    double x, y, z, occ, bf;
    int8_t charge;
};

EMM_HOT uint32_t cif_pack(const CifTok &t, int width)
{
    uint32_t v = 0;
    if (t.p && !t.null())
        for (int i = 0; i < width && i < t.len; ++i) v |= (uint32_t)(unsigned char)t.p[i] << (8 * i);
    return v;
}

EMM_HOT bool cif_real(const CifTok &t, double *out, bool optional)
{
    if (!t.p || t.null() || t.len == 0) { *out = 0.0; return optional; }
    if (fast_real(t.p, t.p + t.len, out)) return true;
    char buf[64];
    if (t.len >= (int64_t)sizeof buf) return false;
    memcpy(buf, t.p, (size_t)t.len);
    buf[t.len] = 0;
    if (char *paren = strchr(buf, '(')) *paren = 0;             // "1.234(5)": a standard uncertainty
    char *stop = nullptr;
    *out = strtod(buf, &stop);
    return stop != buf && *stop == 0;
}

EMM_HOT bool cif_int(const CifTok &t, int32_t *out)
{
    return t.p && !t.null() && t.len > 0 && fast_int(t.p, t.p + t.len, out);
}

// One _atom_site row -> AtomRec; false with t_error set when a required item is missing or malformed.
bool cif_row(const CifTok *v, int64_t row, int flags, AtomRec &a)
{
    const bool author = (flags & EMM_PDB_CIF_AUTHOR) != 0;
    auto pick = [&](int label, int auth) -> const CifTok & {
        const CifTok &first = v[author ? auth : label], &second = v[author ? label : auth];
        return (first.p && !first.null()) ? first : second;
    };
    if (!cif_real(v[C_X], &a.x, false) || !cif_real(v[C_Y], &a.y, false) || !cif_real(v[C_Z], &a.z, false) ||
        !cif_real(v[C_OCC], &a.occ, true) || !cif_real(v[C_B], &a.bf, true)) {
        t_error = "malformed _atom_site row " + std::to_string(row + 1) + ": coordinates, occupancy or B factor";
        return false;
    }
    if (!cif_int(v[C_ID], &a.serial)) a.serial = (int32_t)(row + 1);
    if (!cif_int(pick(C_L_SEQ, C_A_SEQ), &a.resnum)) a.resnum = 0;           // '.' for non-polymers without auth_seq_id
    const CifTok &chain = pick(C_L_ASYM, C_A_ASYM);
    if (chain.p && !chain.null() && chain.len > 2) {
        t_error = "_atom_site row " + std::to_string(row + 1) + ": chain id '" + std::string(chain.p, (size_t)chain.len) +
                  "' is longer than the two characters an Atom.chain_id holds";
        return false;
    }
    a.chain = (uint16_t)cif_pack(chain, 2);
    a.name = cif_pack(pick(C_L_ATOM, C_A_ATOM), 4);
    a.resname = cif_pack(pick(C_L_COMP, C_A_COMP), 4);
    a.element = (uint16_t)cif_pack(v[C_SYMBOL], 2);
    a.altloc = (v[C_ALT].p && !v[C_ALT].null() && v[C_ALT].len > 0) ? v[C_ALT].p[0] : ' ';
    a.icode = (v[C_INS].p && !v[C_INS].null() && v[C_INS].len > 0) ? v[C_INS].p[0] : ' ';
    int32_t q = 0;
    a.charge = (int8_t)(cif_int(v[C_CHARGE], &q) ? q : 0);
    return true;
}

// Calls on_atom(rec) for every _atom_site row of the first model, in file order (the mmCIF counterpart
// of "ATOM and HETATM records up to the first ENDMDL").  Returns the number of atoms delivered or -1
// (t_error set).  header_id receives the data block name when it fits four characters.
template <class F>
int64_t cif_each_atom(const char *text, int64_t len, int flags, char header_id[5], F &&on_atom)
{
    memset(header_id, 0, 5);
    CifScanner scan(text, len);
    CifTok t;
    if (!scan.next(t) || !t.starts("data_")) { t_error = "not an mmCIF text: no data_ block"; return -1; }
    if (t.len - 5 >= 1 && t.len - 5 <= 4) memcpy(header_id, t.p + 5, (size_t)(t.len - 5));
    CifTok row[C_N], pairs[C_N];
    bool have_pairs = false;
    for (int k = 0; k < C_N; ++k) pairs[k] = CifTok{nullptr, 0, false};
    bool more = scan.next(t);
    while (more) {
        if (t.starts("data_")) break;                                  // a second block: not ours
        if (t.is("loop_")) {
            std::vector<int> cols;
            bool ours = false;
            while ((more = scan.next(t)) && !t.quoted && t.len > 0 && t.p[0] == '_') {
                if (cols.empty()) ours = t.starts("_atom_site.");
                cols.push_back(ours ? cif_column(t) : -1);
            }
            if (!ours) continue;                                       // its values are skipped by the main loop
            if (cols.empty()) { t_error = "_atom_site loop without items"; return -1; }
            bool have[C_N] = {false};
            for (int c : cols) if (c >= 0) have[c] = true;
            if (!have[C_X] || !have[C_Y] || !have[C_Z]) { t_error = "_atom_site loop without Cartn_x / Cartn_y / Cartn_z"; return -1; }
            int64_t n = 0;
            CifTok model0{nullptr, 0, false};
            const size_t ncols = cols.size();
            for (int k = 0; k < C_N; ++k) row[k] = CifTok{nullptr, 0, false};      // mapped entries are rewritten by every row
            while (more) {
                // a row starts unless the token ends the loop: a tag or one of the reserved words
                if (!t.quoted && t.len > 0) {
                    const char c0 = t.p[0];
                    if (c0 == '_') break;
                    if ((c0 == 'l' || c0 == 'L' || c0 == 'd' || c0 == 'D' || c0 == 's' || c0 == 'S' || c0 == 'g' || c0 == 'G') &&
                        (t.is("loop_") || t.starts("data_") || t.starts("save_") || t.is("stop_") || t.is("global_")))
                        break;
                }
                for (size_t j = 0; j < ncols; ++j) {
                    if (!more) { t_error = "_atom_site loop ends in the middle of a row"; return -1; }
                    if (cols[j] >= 0) row[cols[j]] = t;
                    more = scan.next(t);
                }
                if (row[C_MODEL].p) {
                    if (!model0.p) model0 = row[C_MODEL];
                    else if (model0.len != row[C_MODEL].len || memcmp(model0.p, row[C_MODEL].p, (size_t)model0.len) != 0) return n;
                }
                AtomRec a;
                if (!cif_row(row, n, flags, a)) return -1;
                if (!on_atom(a)) return -1;
                ++n;
            }
            return n;
        }
        if (t.starts("_atom_site.")) {                                 // item-value pairs: a one-atom category
            const int c = cif_column(t);
            more = scan.next(t);
            if (!more) break;
            if (c >= 0) { pairs[c] = t; have_pairs = true; }
        }
        more = scan.next(t);
    }
    if (have_pairs) {
        if (!pairs[C_X].p || !pairs[C_Y].p || !pairs[C_Z].p) { t_error = "_atom_site without Cartn_x / Cartn_y / Cartn_z"; return -1; }
        AtomRec a;
        if (!cif_row(pairs, 0, flags, a) || !on_atom(a)) return -1;
        return 1;
    }
    return 0;
}

inline void unpack_field(uint32_t v, char *dst, int width)
{
    for (int i = 0; i < width; ++i) dst[i] = (char)((v >> (8 * i)) & 0xffu);
}

inline void cif_store(const AtomRec &a, const Columns &c, int64_t i)
{
    c.serial[i] = a.serial; c.resnum[i] = a.resnum;
    c.xyz[3 * i] = a.x; c.xyz[3 * i + 1] = a.y; c.xyz[3 * i + 2] = a.z;
    c.occupancy[i] = a.occ; c.bfactor[i] = a.bf;
    unpack_field(a.name, c.name + 4 * i, 4);
    unpack_field(a.resname, c.resname + 4 * i, 4);
    unpack_field(a.chain, c.chain + 2 * i, 2);
    unpack_field(a.element, c.element + 2 * i, 2);
    memset(c.segment + 4 * i, 0, 4);
    c.altloc[i] = a.altloc; c.icode[i] = a.icode; c.charge[i] = a.charge;
}

int64_t cif_parse_into(const char *text, int64_t len, int flags, const Columns &c, int64_t base, int64_t capacity, char header_id[5])
{
    int64_t n = 0;
    return cif_each_atom(text, len, flags, header_id, [&](const AtomRec &a) {
        if (n >= capacity) { t_error = "atom capacity exceeded"; return false; }
        cif_store(a, c, base + n++);
        return true;
    });
}

// The readers that take paths tokenise an mmCIF text ONCE: the rows are collected here, counted, and
// the columns filled from the records (the text API counts and parses in two calls, so it reads twice).
int64_t cif_collect(const char *text, int64_t len, int flags, char header_id[5], std::vector<AtomRec> &atoms)
{
    atoms.clear();
    return cif_each_atom(text, len, flags, header_id, [&](const AtomRec &a) { atoms.push_back(a); return true; });
}

void cif_fill_packed(const std::vector<AtomRec> &atoms, const PackedCols &c, KindTable &kinds, std::vector<uint64_t> &run_keys,
                     bool *split, int32_t *residue_count)
{
    std::vector<int32_t> counted;
    PackState state(c, kinds, run_keys, counted);
    for (size_t i = 0; i < atoms.size(); ++i) {
        const AtomRec &a = atoms[i];
        c.xyz[3 * i] = a.x; c.xyz[3 * i + 1] = a.y; c.xyz[3 * i + 2] = a.z;
        state.add((int64_t)i, a.name, a.resname, a.chain, a.resnum, (float)a.bf);
    }
    state.finish(split, residue_count);
}

// ---- either format --------------------------------------------------------------------------------

int64_t count_atoms(const char *text, int64_t len, int flags)
{
    if (!looks_like_cif(text, len)) return pdb_count_atoms(text, len);
    char id[5];
    return cif_each_atom(text, len, flags, id, [](const AtomRec &) { return true; });
}

int64_t parse_into(const char *text, int64_t len, int flags, const Columns &c, int64_t base, int64_t capacity, char header_id[5])
{
    return looks_like_cif(text, len) ? cif_parse_into(text, len, flags, c, base, capacity, header_id)
                                     : pdb_parse_into(text, len, c, base, capacity, header_id);
}

// ---- gzip -----------------------------------------------------------------------------------------

inline bool is_gzip(const char *p, int64_t len) { return len >= 2 && (unsigned char)p[0] == 0x1f && (unsigned char)p[1] == 0x8b; }

// zlib is looked up at run time (the library does not link it: only gzip input needs it)
struct Zlib {
    int (*init2)(z_streamp, int, const char *, int) = nullptr;
    int (*run)(z_streamp, int) = nullptr;
    int (*reset)(z_streamp) = nullptr;
    int (*end)(z_streamp) = nullptr;
    bool ok = false;
    Zlib()
    {
        void *h = dlopen("libz.so.1", RTLD_NOW | RTLD_LOCAL);
        if (!h) h = dlopen("libz.so", RTLD_NOW | RTLD_LOCAL);
        if (!h) return;
        init2 = reinterpret_cast<int (*)(z_streamp, int, const char *, int)>(dlsym(h, "inflateInit2_"));
        run = reinterpret_cast<int (*)(z_streamp, int)>(dlsym(h, "inflate"));
        reset = reinterpret_cast<int (*)(z_streamp)>(dlsym(h, "inflateReset"));
        end = reinterpret_cast<int (*)(z_streamp)>(dlsym(h, "inflateEnd"));
        ok = init2 && run && reset && end;
    }
};

// inflate a (possibly multi-member) gzip image into a reusable buffer
bool gunzip(const char *src, int64_t n, std::vector<char> &dst, int64_t *len)
{
    static const Zlib z;
    if (!z.ok) return false;
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (z.init2(&zs, 16 + MAX_WBITS, ZLIB_VERSION, (int)sizeof(z_stream)) != Z_OK) return false;
    if (dst.size() < (size_t)n * 4 + 65536) dst.resize((size_t)n * 4 + 65536);
    size_t got = 0, fed = 0;
    int rc = Z_OK;
    for (;;) {
        if (got == dst.size()) dst.resize(dst.size() * 2);
        const size_t in_now = std::min<size_t>((size_t)n - fed, 1u << 30), out_now = std::min<size_t>(dst.size() - got, 1u << 30);
        zs.next_in = reinterpret_cast<Bytef *>(const_cast<char *>(src)) + fed;
        zs.avail_in = (uInt)in_now;
        zs.next_out = reinterpret_cast<Bytef *>(dst.data()) + got;
        zs.avail_out = (uInt)out_now;
        rc = z.run(&zs, Z_NO_FLUSH);
        fed += in_now - zs.avail_in;
        got += out_now - zs.avail_out;
        if (rc == Z_STREAM_END) {
            if (fed >= (size_t)n || !is_gzip(src + fed, n - (int64_t)fed)) break;      // trailing garbage is ignored, as gzip does
            if (z.reset(&zs) != Z_OK) { rc = Z_DATA_ERROR; break; }
            continue;
        }
        if (rc != Z_OK && rc != Z_BUF_ERROR) break;
        if (rc == Z_BUF_ERROR && fed >= (size_t)n && got < dst.size()) break;           // truncated stream
    }
    z.end(&zs);
    if (rc != Z_STREAM_END) return false;
    *len = (int64_t)got;
    return true;
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
    if (is_gzip(out.data(), (int64_t)out.size())) {
        std::vector<char> plain;
        int64_t len = 0;
        if (!gunzip(out.data(), (int64_t)out.size(), plain, &len)) { *err = 4; return false; }
        out.assign(plain.data(), (size_t)len);
    }
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

// message for a per-file error code: 1 open, 2 read, 3 parse (detail), 4 inflate
std::string file_error(int code, const char *path, const std::string &detail)
{
    switch (code) {
    case 1: return std::string("cannot open ") + path;
    case 3: return std::string(path) + ": " + detail;
    case 4: return std::string("cannot inflate ") + path + ": not a complete gzip stream (or zlib is not installed)";
    default: return std::string("cannot read ") + path;
    }
}

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
    // EMM_PDB_SKIP_BAD: what went wrong with the files that were skipped (0 = fine; file_error codes)
    std::vector<int32_t> file_status;
    std::vector<std::string> file_message;
};

extern "C" {

const char *emm_pdb_last_error(void) { return t_error.c_str(); }

int emm_pdb_count_atoms(const char *text, int64_t len, int64_t *n_atoms)
{
    if (!text || !n_atoms || len < 0) return EMM_ERR_INVALID;
    const int64_t n = count_atoms(text, len, 0);
    if (n < 0) return EMM_ERR_INPUT;
    *n_atoms = n;
    return EMM_OK;
}

int emm_pdb_parse_ex(const char *text, int64_t len, int32_t flags, int64_t capacity, int32_t *serial, char *name,
                     char *altloc, char *resname, char *chain, int32_t *resnum, char *icode, double *xyz,
                     double *occupancy, double *bfactor, char *segment, char *element, int8_t *charge, char *header_id,
                     int64_t *n_atoms)
{
    if (!text || !n_atoms || len < 0 || !header_id) return EMM_ERR_INVALID;
    Columns c{serial, name, altloc, resname, chain, resnum, icode, xyz, occupancy, bfactor, segment, element, charge};
    const int64_t n = parse_into(text, len, flags, c, 0, capacity, header_id);
    if (n < 0) return EMM_ERR_INPUT;
    *n_atoms = n;
    return EMM_OK;
}

int emm_pdb_parse(const char *text, int64_t len, int64_t capacity, int32_t *serial, char *name, char *altloc,
                  char *resname, char *chain, int32_t *resnum, char *icode, double *xyz, double *occupancy,
                  double *bfactor, char *segment, char *element, int8_t *charge, char *header_id, int64_t *n_atoms)
{
    return emm_pdb_parse_ex(text, len, 0, capacity, serial, name, altloc, resname, chain, resnum, icode, xyz, occupancy,
                            bfactor, segment, element, charge, header_id, n_atoms);
}

int emm_pdb_load_files(const char *const *paths, int32_t n_files, int32_t n_threads, emm_pdb_batch **out)
{
    return emm_pdb_load_files_ex(paths, n_files, n_threads, 0, out);
}

int emm_pdb_load_files_ex(const char *const *paths, int32_t n_files, int32_t n_threads, int32_t flags, emm_pdb_batch **out)
{
    if (!paths || !out || n_files < 0) return EMM_ERR_INVALID;
    *out = nullptr;
    emm_pdb_batch *b = new emm_pdb_batch();
    b->n_files = n_files;
    std::vector<std::string> texts((size_t)n_files);
    std::vector<int64_t> counts((size_t)n_files, 0);
    std::vector<int> errs((size_t)n_files, 0);
    std::vector<std::string> messages((size_t)n_files);
    std::vector<char> is_cif((size_t)n_files, 0);
    std::vector<std::vector<AtomRec>> cif_atoms((size_t)n_files);
    b->header_id.assign(5 * (size_t)n_files, 0);
    if (n_threads < 1) n_threads = 1;
    n_threads = std::min<int32_t>(n_threads, std::max(n_files, 1));
    {
        std::atomic<int> next(0);
        auto work = [&]() {
            for (int i; (i = next.fetch_add(1)) < n_files;) {
                if (read_file(paths[i], texts[(size_t)i], &errs[(size_t)i])) {
                    std::string &text = texts[(size_t)i];
                    if (looks_like_cif(text.data(), (int64_t)text.size())) {
                        is_cif[(size_t)i] = 1;
                        counts[(size_t)i] = cif_collect(text.data(), (int64_t)text.size(), flags, &b->header_id[5 * (size_t)i],
                                                        cif_atoms[(size_t)i]);
                        std::string().swap(text);                       // the records are all pass 2 needs
                    } else {
                        counts[(size_t)i] = pdb_count_atoms(text.data(), (int64_t)text.size());
                    }
                    if (counts[(size_t)i] < 0) { counts[(size_t)i] = 0; errs[(size_t)i] = 3; messages[(size_t)i] = t_error; }
                }
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
        work();
        for (auto &t : pool) t.join();
    }
    for (int i = 0; i < n_files; ++i)
        if (errs[(size_t)i]) {
            t_error = file_error(errs[(size_t)i], paths[i], messages[(size_t)i]);
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
    Columns c{b->serial.data(), b->name.data(), b->altloc.data(), b->resname.data(), b->chain.data(), b->resnum.data(),
              b->icode.data(), b->xyz.data(), b->occupancy.data(), b->bfactor.data(), b->segment.data(),
              b->element.data(), b->charge.data()};
    std::atomic<int> next(0), failed(-1);
    auto work = [&]() {
        for (int i; (i = next.fetch_add(1)) < n_files;) {
            if (is_cif[(size_t)i]) {
                std::vector<AtomRec> &atoms = cif_atoms[(size_t)i];
                for (size_t a = 0; a < atoms.size(); ++a) cif_store(atoms[a], c, b->atom_off[(size_t)i] + (int64_t)a);
                std::vector<AtomRec>().swap(atoms);
                continue;
            }
            const int64_t got = pdb_parse_into(texts[(size_t)i].data(), (int64_t)texts[(size_t)i].size(), c,
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
    return emm_pdb_pack_files_ex(paths, n_files, n_threads, 0, out);
}

int emm_pdb_pack_files_ex(const char *const *paths, int32_t n_files, int32_t n_threads, int32_t flags, emm_pdb_batch **out)
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
            std::vector<char> text, inflated;
            std::vector<AtomRec> cif_atoms;
            for (int i; (i = next.fetch_add(1)) < n_files;) {
                const size_t f = (size_t)i;
                int64_t len = 0;
                if (!read_file_into(paths[i], text, &len, &errs[f])) continue;
                const char *data = text.data();
                if (is_gzip(data, len)) {
                    if (!gunzip(data, len, inflated, &len)) { errs[f] = 4; continue; }
                    data = inflated.data();
                }
                FileBlock &blk = blocks[f];
                const bool cif = looks_like_cif(data, len);
                const int64_t count = cif ? cif_collect(data, len, flags, &b->header_id[5 * f], cif_atoms) : pdb_count_atoms(data, len);
                if (count < 0) { messages[f] = t_error; errs[f] = 3; continue; }
                blk.allocate(count);
                kinds.clear();
                const PackedCols c{blk.xyz, blk.kind, blk.residue, blk.bfactor, blk.chain};
                if (cif) {
                    cif_fill_packed(cif_atoms, c, kinds, run_keys, &blk.split, &blk.residue_count);
                } else {
                    const int64_t got = pdb_pack_into(data, len, c, 0, count, kinds, run_keys, &blk.split, &b->header_id[5 * f],
                                                      &blk.residue_count);
                    if (got != count) { messages[f] = t_error; errs[f] = 3; continue; }
                }
                blk.kinds = kinds.keys;
                if (blk.split) {
                    blk.atom_id.reset(new int32_t[(size_t)std::max<int64_t>(count, 1)]);
                    regroup_file(c, 0, count, run_keys, blk.atom_id.get());
                }
                blk.res_keys = run_keys;
            }
        });
    }
    b->file_status.assign(nf, 0);
    b->file_message.assign(nf, std::string());
    for (int i = 0; i < n_files; ++i)
        if (errs[(size_t)i]) {
            const int e = errs[(size_t)i];
            t_error = file_error(e, paths[i], messages[(size_t)i]);
            if (!(flags & EMM_PDB_SKIP_BAD)) return e == 1 ? EMM_ERR_INVALID : EMM_ERR_INPUT;
            // skipped and named: the file stays in the batch as a structure without atoms
            b->file_status[(size_t)i] = e;
            b->file_message[(size_t)i] = t_error;
            blocks[(size_t)i] = FileBlock();
            blocks[(size_t)i].allocate(0);
            memset(&b->header_id[5 * (size_t)i], 0, 5);
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
                const PackedCols c{blk.xyz, blk.kind, blk.residue, blk.bfactor, blk.chain};
                std::vector<int32_t> counted;
                PackState state(c, kinds, run_keys, counted);
                const uint8_t *nm = name4[f], *rn = resname4[f], *ch = chain2[f];
                for (int64_t a = 0; a < n; ++a) {
                    uint32_t name, res;
                    uint16_t chain;
                    memcpy(&name, nm + 4 * a, 4);
                    memcpy(&res, rn + 4 * a, 4);
                    memcpy(&chain, ch + 2 * a, 2);
                    state.add(a, name, res, chain, resnum[f][a], (float)bfactor[f][a]);
                }
                if (n) memcpy(blk.xyz, xyz[f], (size_t)n * 3 * sizeof(double));
                blk.kinds = kinds.keys;
                state.finish(&blk.split, &blk.residue_count);
                if (blk.split) {
                    blk.atom_id.reset(new int32_t[(size_t)std::max<int64_t>(n, 1)]);
                    regroup_file(c, 0, n, run_keys, blk.atom_id.get());
                }
                blk.res_keys = run_keys;
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

int emm_pdb_batch_file_status(const emm_pdb_batch *b, int32_t *status, int32_t capacity)
{
    if (!b || !status || capacity < b->n_files) return EMM_ERR_INVALID;
    for (int32_t i = 0; i < b->n_files; ++i) status[i] = (size_t)i < b->file_status.size() ? b->file_status[(size_t)i] : 0;
    return EMM_OK;
}

const char *emm_pdb_batch_file_message(const emm_pdb_batch *b, int32_t file)
{
    if (!b || file < 0 || (size_t)file >= b->file_message.size()) return "";
    return b->file_message[(size_t)file].c_str();
}

void emm_pdb_batch_free(emm_pdb_batch *b) { delete b; }

}  // extern "C"
