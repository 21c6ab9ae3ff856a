/* Plain-C host for the C-ABI of libchessvision_b200.so: no torch, no Python, no C++.
 *
 *   gcc -std=c99 -I include examples/c_abi_demo.c -o c_abi_demo -L chess_vision_b200 -lchessvision_b200 \
 *       -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,'$ORIGIN/../chess_vision_b200'
 *   ./c_abi_demo weights.cvb|weights.cvs [n_boards]
 *
 * weights.cvb is the packed-weights file chess_vision_b200.checkpoint.save_packed() writes from a reference checkpoint
 * (MAGIC "CVB200W1" | u32 header bytes | header JSON | fp32 blob); weights.cvs is the RAW state_dict
 * (checkpoint.save_raw_state_dict: MAGIC "CVB200S1" | u32 count | per tensor u16 name bytes | name | i64 numel | fp32 data), which
 * this program packs itself through cv_square_pack_weights -- BatchNorm fold included, no Python anywhere.  It then generates
 * synthetic boards with the library's own counter-based generator, runs the host-buffer entry point (the call predict.py's user
 * makes, batched) and prints the FEN strings -- the same bytes the Python surface returns (tests/test_c_abi_demo.py). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "chessvision_b200.h"

/* the only CUDA runtime calls a host needs: device memory for the weight blob */
extern int cudaMalloc(void** p, size_t n);
extern int cudaFree(void* p);
extern int cudaMemcpy(void* dst, const void* src, size_t n, int kind);   /* kind 1 = host to device */

#define CHECK(call)                                                            \
    do {                                                                       \
        int rc_ = (call);                                                      \
        if (rc_ != 0) { fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, cv_last_error()); return 1; } \
    } while (0)

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s weights.cvb [n_boards]\n", argv[0]); return 2; }
    const int B = argc > 2 ? atoi(argv[2]) : 4, H = 256;
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    char magic[8];
    const size_t n = cv_weight_blob_floats();
    float* blob = (float*)malloc(n * sizeof(float));
    if (fread(magic, 1, 8, f) != 8) { fprintf(stderr, "not a weights file\n"); return 2; }
    if (memcmp(magic, "CVB200W1", 8) == 0) {                                 /* packed blob */
        uint32_t hdr = 0;
        if (fread(&hdr, 4, 1, f) != 1) { fprintf(stderr, "truncated header\n"); return 2; }
        fseek(f, (long)hdr, SEEK_CUR);
        if (fread(blob, sizeof(float), n, f) != n) { fprintf(stderr, "truncated weight blob\n"); return 2; }
    } else if (memcmp(magic, "CVB200S1", 8) == 0) {                          /* raw state_dict: pack it here */
        uint32_t count = 0;
        if (fread(&count, 4, 1, f) != 1 || count > 4096) { fprintf(stderr, "bad tensor count\n"); return 2; }
        cv_named_tensor* t = (cv_named_tensor*)calloc(count, sizeof(cv_named_tensor));
        for (uint32_t i = 0; i < count; ++i) {
            uint16_t len = 0;
            int64_t numel = 0;
            if (fread(&len, 2, 1, f) != 1) { fprintf(stderr, "truncated state_dict\n"); return 2; }
            char* name = (char*)calloc((size_t)len + 1, 1);
            if (fread(name, 1, len, f) != len || fread(&numel, 8, 1, f) != 1 || numel < 0) { fprintf(stderr, "truncated state_dict\n"); return 2; }
            float* data = (float*)malloc((size_t)(numel ? numel : 1) * sizeof(float));
            if (fread(data, sizeof(float), (size_t)numel, f) != (size_t)numel) { fprintf(stderr, "truncated tensor %s\n", name); return 2; }
            t[i].name = name; t[i].data = data; t[i].numel = numel;
        }
        CHECK(cv_square_pack_weights(t, (int)count, blob, n));
        for (uint32_t i = 0; i < count; ++i) { free((void*)t[i].name); free((void*)t[i].data); }
        free(t);
    } else { fprintf(stderr, "not a chess_vision_b200 weights file\n"); return 2; }
    fclose(f);

    cv_square* h = NULL;
    CHECK(cv_square_create(0, &h));
    void* blob_dev = NULL;
    if (cudaMalloc(&blob_dev, n * sizeof(float)) != 0 || cudaMemcpy(blob_dev, blob, n * sizeof(float), 1) != 0) { fprintf(stderr, "cudaMalloc/cudaMemcpy failed\n"); return 1; }
    CHECK(cv_square_load_weights(h, (const float*)blob_dev, n, NULL));

    uint8_t* boards = (uint8_t*)malloc((size_t)B * H * H * 3);
    char* fen = (char*)calloc((size_t)B, CV_FEN_STRIDE);
    uint8_t* fen_len = (uint8_t*)calloc((size_t)B, 1);
    CHECK(cv_synth_boards_host(boards, CV_LAYOUT_HWC, 0, B, H, 1u, 1, NULL));            /* structured boards, seed 1 */
    CHECK(cv_square_predict_host_u8(h, boards, CV_LAYOUT_HWC, NULL, B, H, CV_PRECISION_FP32, fen, fen_len));
    for (int b = 0; b < B; ++b) printf("%.*s\n", (int)fen_len[b], fen + (size_t)b * CV_FEN_STRIDE);

    CHECK(cv_square_destroy(h));
    cudaFree(blob_dev);
    free(blob); free(boards); free(fen); free(fen_len);
    return 0;
}
