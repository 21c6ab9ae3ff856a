/* Plain-C host for the C-ABI of libchessvision_b200.so: no torch, no Python, no C++.
 *
 *   gcc -std=c99 -I include examples/c_abi_demo.c -o c_abi_demo -L chess_vision_b200 -lchessvision_b200 \
 *       -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,'$ORIGIN/../chess_vision_b200'
 *   ./c_abi_demo weights.cvb [n_boards]
 *
 * weights.cvb is the packed-weights file chess_vision_b200.checkpoint.save_packed() writes from a reference checkpoint
 * (MAGIC "CVB200W1" | u32 header bytes | header JSON | fp32 blob).  The program generates synthetic boards with the library's
 * own counter-based generator, runs the host-buffer entry point (the call predict.py's user makes, batched) and prints the
 * FEN strings -- the same bytes the Python surface returns (tests/test_c_abi_demo.py). */
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
    uint32_t hdr = 0;
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "CVB200W1", 8) != 0 || fread(&hdr, 4, 1, f) != 1) { fprintf(stderr, "not a packed-weights file\n"); return 2; }
    fseek(f, (long)hdr, SEEK_CUR);
    const size_t n = cv_weight_blob_floats();
    float* blob = (float*)malloc(n * sizeof(float));
    if (fread(blob, sizeof(float), n, f) != n) { fprintf(stderr, "truncated weight blob\n"); return 2; }
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
