// emm_tsv.cpp -- native writer for the rows of EnzyMM's results table (host only; needs no GPU).
//
// Replaces the per-match Python of Match.dump (reference enzymm/jess_run.py:185-284) for screening
// runs: at > 10^4 structures/s the table is tens of thousands of rows per second, and building a
// Match, its Atom objects and a csv.writer row per hit costs more than the search.  The caller
// (enzymm_b200/tsv.py) decides WHICH hits become rows and in which order -- size groups,
// completeness, the filter verdict, match indices: Matcher semantics, vectorised on the host -- and
// gathers the few per-atom fields a row needs; this file only formats, byte for byte what Python
// prints: str(round(x, 5)) for the three floats, str(bool), the csv module's minimal quoting.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/enzymm_b200.h"

namespace {

// repr(float) of Python 3: the shortest digit string that round-trips, in fixed notation when the
// decimal exponent is in [-4, 16), otherwise d.ddde-XX.
void py_repr(double v, std::string &out)
{
    if (std::isnan(v)) { out += "nan"; return; }
    if (std::isinf(v)) { out += v < 0 ? "-inf" : "inf"; return; }
    char buf[40];
    int prec = 0;
    for (prec = 0; prec < 17; ++prec) {           // prec + 1 significant digits
        snprintf(buf, sizeof buf, "%.*e", prec, v);
        if (strtod(buf, nullptr) == v) break;
    }
    // buf = [-]d[.ddd]e[+-]XX
    const char *p = buf;
    if (*p == '-') { out += '-'; ++p; }
    std::string digits;
    digits += *p++;
    if (*p == '.') { ++p; while (*p != 'e') digits += *p++; }
    const int exp10 = atoi(p + 1);
    while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
    const int nd = (int)digits.size();
    if (exp10 < -4 || exp10 >= 16) {
        out += digits[0];
        if (nd > 1) { out += '.'; out.append(digits, 1, std::string::npos); }
        char e[16];
        snprintf(e, sizeof e, "e%c%02d", exp10 < 0 ? '-' : '+', std::abs(exp10));
        out += e;
    } else if (exp10 < 0) {
        out += "0.";
        out.append((size_t)(-exp10 - 1), '0');
        out += digits;
    } else {
        if (nd <= exp10 + 1) {
            out += digits;
            out.append((size_t)(exp10 + 1 - nd), '0');
            out += ".0";
        } else {
            out.append(digits, 0, (size_t)exp10 + 1);
            out += '.';
            out.append(digits, (size_t)exp10 + 1, std::string::npos);
        }
    }
}

// str(round(v, 5)): Python rounds through the correctly rounded 5-decimal string, as printf does
void py_round5(double v, std::string &out)
{
    if (std::isnan(v) || std::isinf(v)) { py_repr(v, out); return; }
    char buf[400];
    snprintf(buf, sizeof buf, "%.5f", v);
    py_repr(strtod(buf, nullptr), out);
}

// csv.writer with QUOTE_MINIMAL, delimiter '\t', quotechar '"', lineterminator '\n'
void csv_field(const char *s, std::string &out)
{
    bool quote = false;
    for (const char *p = s; *p; ++p)
        if (*p == '\t' || *p == '"' || *p == '\n' || *p == '\r') { quote = true; break; }
    if (!quote) { out += s; return; }
    out += '"';
    for (const char *p = s; *p; ++p) { if (*p == '"') out += '"'; out += *p; }
    out += '"';
}

inline void fixed_field(const char *p, int width, std::string &out)
{
    for (int i = 0; i < width && p[i]; ++i) out += p[i];
}

}  // namespace

extern "C" {

int emm_tsv_format(const emm_tsv_rows *r, char **text, int64_t *len)
{
    if (!r || !text || !len || r->n_rows < 0) return EMM_ERR_INVALID;
    std::string out;
    out.reserve((size_t)r->n_rows * 256 + 16);
    char num[32];
    for (int64_t i = 0; i < r->n_rows; ++i) {
        const int t = r->template_index[i], s = r->structure[i];
        const int n = r->n_atoms[i];
        if (n < 0 || n > EMM_MAX_TEMPLATE_ATOMS) return EMM_ERR_INVALID;
        const char *resname = r->resname4 + (size_t)i * EMM_MAX_TEMPLATE_ATOMS * 4;
        const char *chain = r->chain2 + (size_t)i * EMM_MAX_TEMPLATE_ATOMS * 2;
        const int32_t *resnum = r->resnum + (size_t)i * EMM_MAX_TEMPLATE_ATOMS;
        csv_field(r->query_id[s], out); out += '\t';
        out += r->tpl_distance[t]; out += '\t';
        snprintf(num, sizeof num, "%d", r->match_index[i]); out += num; out += '\t';
        out += r->tpl_static[t]; out += '\t';
        out += r->tpl_multimeric[t] ? "True" : "False"; out += '\t';
        bool multimeric = false;                                   // Match.multimeric, jess_run.py:375-383
        for (int a = 1; a < n; ++a) multimeric = multimeric || memcmp(chain + 2 * a, chain, 2) != 0;
        out += multimeric ? "True" : "False"; out += '\t';
        snprintf(num, sizeof num, "%d", r->query_atom_count[s]); out += num; out += '\t';
        snprintf(num, sizeof num, "%d", r->query_residue_count[s]); out += num; out += '\t';
        py_round5(r->rmsd[i], out); out += '\t';
        py_round5(r->log_evalue[i], out); out += '\t';
        py_round5(r->orientation[i], out); out += '\t';
        // Match.preserved_resid_order (jess_run.py:385-423): dense ranks of the matched residue numbers
        const int n_res = n / 3;
        bool preserved = !(r->tpl_multimeric[t] || multimeric);
        if (preserved) {
            const int o0 = r->tpl_order_off[t];
            if (r->tpl_order_off[t + 1] - o0 != n_res) preserved = false;
            for (int j = 0; j < n_res && preserved; ++j) {
                int rank = 1;                                      // 1 + number of distinct smaller values
                for (int q = 0; q < n_res; ++q) {
                    if (resnum[3 * q] >= resnum[3 * j]) continue;
                    bool seen = false;
                    for (int p = 0; p < q; ++p) seen = seen || resnum[3 * p] == resnum[3 * q];
                    rank += seen ? 0 : 1;
                }
                preserved = rank == r->tpl_order[o0 + j];
            }
        }
        out += preserved ? "True" : "False"; out += '\t';
        out += r->complete[i] ? "True" : "False"; out += '\t';
        if (r->predicted[i] < 2) out += r->predicted[i] ? "True" : "False";
        out += '\t';
        {
            std::string residues;                                  // RES_chain_number per atom triplet
            for (int j = 0; j < n_res; ++j) {
                if (j) residues += ',';
                fixed_field(resname + 4 * (3 * j), 4, residues); residues += '_';
                fixed_field(chain + 2 * (3 * j), 2, residues); residues += '_';
                snprintf(num, sizeof num, "%d", resnum[3 * j]); residues += num;
            }
            csv_field(residues.c_str(), out);
        }
        out += '\t';
        out += r->tpl_annotation[t];
        out += '\n';
    }
    char *mem = (char *)malloc(out.size() + 1);
    if (!mem) return EMM_ERR_NOMEM;
    memcpy(mem, out.data(), out.size());
    mem[out.size()] = 0;
    *text = mem;
    *len = (int64_t)out.size();
    return EMM_OK;
}

void emm_tsv_free(char *text) { free(text); }

int emm_tsv_repr_round5(double v, char *out, int32_t capacity)
{
    std::string s;
    py_round5(v, s);
    if (!out || capacity <= (int32_t)s.size()) return EMM_ERR_INVALID;
    memcpy(out, s.c_str(), s.size() + 1);
    return EMM_OK;
}

}  // extern "C"
