/* A plain C99 client of include/enzymm_b200.h, the way a cgo / JNI / ctypes-free binding would use
 * it: sanity constants, the loud no-device error, and the host-only PDB reader on argv[1].
 * Built and run by tests/test_cabi.py::test_c99_client_links_and_runs (no GPU needed). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "enzymm_b200.h"

int main(int argc, char **argv)
{
    if (emm_abi_version() < 2) { fprintf(stderr, "abi %d\n", emm_abi_version()); return 1; }
    if (emm_hit_size() != (int)sizeof(emm_hit)) { fprintf(stderr, "emm_hit: header %zu, library %d\n", sizeof(emm_hit), emm_hit_size()); return 2; }
    const int devices = emm_device_count();
    printf("abi=%d hit=%d devices=%d\n", emm_abi_version(), emm_hit_size(), devices);
    if (devices <= 0) {
        /* no CPU fallback: creating a stream (the first thing a session needs) must say so */
        void *stream = NULL;
        const int rc = emm_stream_create(0, &stream);
        if (rc != EMM_ERR_NO_DEVICE || stream != NULL) { fprintf(stderr, "stream_create rc=%d\n", rc); return 3; }
        const char *why = emm_last_error();
        if (why == NULL || strlen(why) == 0) { fprintf(stderr, "no error text\n"); return 4; }
    }
    if (argc > 1) {
        FILE *f = fopen(argv[1], "rb");
        if (!f) { perror(argv[1]); return 5; }
        fseek(f, 0, SEEK_END);
        const long len = ftell(f);
        fseek(f, 0, SEEK_SET);
        char *text = (char *)malloc((size_t)len + 1);
        if (fread(text, 1, (size_t)len, f) != (size_t)len) { fclose(f); return 6; }
        fclose(f);
        int64_t n = 0;
        if (emm_pdb_count_atoms(text, (int64_t)len, &n) != EMM_OK) { fprintf(stderr, "%s\n", emm_pdb_last_error()); return 7; }
        printf("atoms=%lld\n", (long long)n);
        free(text);
    }
    return 0;
}
