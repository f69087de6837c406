// Mutation fuzzer for the PDB / mmCIF readers of csrc/emm_pdb.cpp, meant to be built with
// -fsanitize=address,undefined together with that file: every input is an exact-size heap copy, so
// an over-read of one byte is a report.  usage: fuzz_readers <iterations> <seed files...>
// (tests/test_cif_ingest.py::test_readers_survive_mutated_inputs_under_sanitizers)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>
#include "enzymm_b200.h"

static std::string slurp(const char *path)
{
    FILE *f = fopen(path, "rb");
    if (!f) { perror(path); exit(2); }
    std::string s;
    char buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) s.append(buf, n);
    fclose(f);
    return s;
}

static void run_one(const std::string &text, int flags)
{
    int64_t n = 0;
    // exact-size heap copy so that any overread trips ASan
    char *copy = (char *)malloc(text.size() ? text.size() : 1);
    memcpy(copy, text.data(), text.size());
    const int rc = emm_pdb_count_atoms(copy, (int64_t)text.size(), &n);
    if (rc == 0) {
        const int64_t cap = n > 0 ? n : 1;
        std::vector<int32_t> serial(cap), resnum(cap);
        std::vector<char> name(4 * cap), altloc(cap), resname(4 * cap), chain(2 * cap), icode(cap), segment(4 * cap), element(2 * cap);
        std::vector<double> xyz(3 * cap), occ(cap), bf(cap);
        std::vector<int8_t> charge(cap);
        char header[5];
        int64_t got = 0;
        emm_pdb_parse_ex(copy, (int64_t)text.size(), flags, n, serial.data(), name.data(), altloc.data(), resname.data(),
                         chain.data(), resnum.data(), icode.data(), xyz.data(), occ.data(), bf.data(), segment.data(),
                         element.data(), charge.data(), header, &got);
    }
    free(copy);
}

// the path-taking readers (one tokenisation for mmCIF, gzip inflation): the same text through a file
static void run_file(const std::string &text, int flags, const char *scratch)
{
    FILE *f = fopen(scratch, "wb");
    if (!f) return;
    fwrite(text.data(), 1, text.size(), f);
    fclose(f);
    const char *paths[2] = {scratch, scratch};
    emm_pdb_batch *b = nullptr;
    if (emm_pdb_pack_files_ex(paths, 2, 2, flags | EMM_PDB_SKIP_BAD, &b) == 0) {
        emm_pdb_packed view;
        emm_pdb_batch_packed(b, &view);
        int32_t status[2];
        emm_pdb_batch_file_status(b, status, 2);
        (void)emm_pdb_batch_file_message(b, 0);
        emm_pdb_batch_free(b);
    }
    b = nullptr;
    if (emm_pdb_load_files_ex(paths, 2, 1, flags, &b) == 0) {
        emm_pdb_columns cols;
        emm_pdb_batch_columns(b, &cols);
        emm_pdb_batch_free(b);
    }
}

int main(int argc, char **argv)
{
    std::mt19937_64 rng(12345);
    const int iters = atoi(argv[1]);
    std::vector<std::string> seeds;
    for (int i = 2; i < argc; ++i) seeds.push_back(slurp(argv[i]));
    const std::string scratch = std::string(argv[2]) + ".fuzz.tmp";      // next to the first seed; gzip seeds are mutated as bytes
    long done = 0;
    for (int it = 0; it < iters; ++it) {
        std::string t = seeds[rng() % seeds.size()];
        // keep it small: a random window that includes the start (so the format is detected) or not
        const bool gz = t.size() > 2 && (unsigned char)t[0] == 0x1f && (unsigned char)t[1] == 0x8b;
        if (!gz && t.size() > 6000) {
            const size_t keep = 500 + rng() % 5500;
            if (rng() % 3) t = t.substr(0, keep);
            else { const size_t off = rng() % (t.size() - keep); t = t.substr(0, 200) + t.substr(off, keep); }
        }
        const int muts = 1 + (int)(rng() % 8);
        for (int m = 0; m < muts && !t.empty(); ++m) {
            const size_t pos = rng() % t.size();
            switch (rng() % 7) {
            case 0: t[pos] = (char)(rng() % 256); break;
            case 1: t.erase(pos, 1 + rng() % 20); break;
            case 2: t.insert(pos, 1 + rng() % 5, "\n;'\"#_ .?"[rng() % 10]); break;
            case 3: t.resize(pos); break;
            case 4: t.insert(pos, "loop_\n"); break;
            case 5: t.insert(pos, "\n_atom_site.Cartn_x "); break;
            case 6: t.insert(pos, "data_x\n"); break;
            }
        }
        run_one(t, (int)(rng() % 2));
        if (it % 8 == 0) run_file(t, (int)(rng() % 2), scratch.c_str());
        ++done;
    }
    remove(scratch.c_str());
    printf("fuzzed %ld inputs\n", done);
    return 0;
}
