/* tests/c/gram_multi_test.c -- a plain C caller of include/stemk.h driving several devices from one host thread:
 * stemk_upload_multi (compile once, device-to-device copies) + stemk_gram_multi against stemk_gram on one device,
 * and the stemk_set_export / stemk_set_import round trip.  Input: a flattened record set written by
 * tests/test_multi_device.py (seven uint64 counts, then the arrays of stemk_seqset_desc in declaration order).
 * Usage: gram_multi_test desc.bin n_devices      prints "GRAM_MULTI OK ..." and exits 0 on success. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "stemk.h"

/* cudart entry points used for the export buffer (the test links libcudart through libstemk_b200.so) */
extern int cudaMalloc(void** p, size_t n);
extern int cudaFree(void* p);
extern int cudaSetDevice(int d);

static void* rd(FILE* f, size_t bytes) {
  void* p = malloc(bytes ? bytes : 1);
  if (!p || (bytes && fread(p, 1, bytes, f) != bytes)) { fprintf(stderr, "short read\n"); exit(2); }
  return p;
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s desc.bin n_devices\n", argv[0]); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  uint64_t c[7];   /* n_seqs, n_nodes, n_edges, n_bpf, n_roots, n_cols, n_weights */
  if (fread(c, sizeof(uint64_t), 7, f) != 7) { fprintf(stderr, "bad header\n"); return 2; }
  const size_t ns = c[0], nn = c[1], ne = c[2], nb = c[3], nr = c[4], nc = c[5], nw = c[6];
  stemk_seqset_desc d;
  d.n_seqs = (uint32_t)ns;
  d.node_off = rd(f, 4 * (ns + 1)); d.node_first = rd(f, 4 * nn); d.node_last = rd(f, 4 * nn); d.node_weight = rd(f, 4 * nn);
  d.edge_off = rd(f, 4 * (nn + 1)); d.edge_to = rd(f, 4 * ne); d.edge_gaps = rd(f, 4 * ne); d.edge_weight = rd(f, 4 * ne);
  d.bpf_off = rd(f, 4 * (nn + 1)); d.bpf_a = rd(f, nb); d.bpf_b = rd(f, nb); d.bpf_freq = rd(f, 4 * nb);
  d.root_off = rd(f, 4 * (ns + 1)); d.root = rd(f, 4 * nr);
  d.col_off = rd(f, 4 * (ns + 1)); d.profile = rd(f, 4 * 5 * nc); d.n_rows = rd(f, 4 * ns);
  d.weight_off = rd(f, 4 * (ns + 1)); d.col_weight = rd(f, 4 * nw); d.text = rd(f, nc);
  fclose(f);

  int n_dev = atoi(argv[2]);
  const int have = stemk_device_count();
  if (n_dev > have) n_dev = have;
  if (n_dev < 1) { fprintf(stderr, "no CUDA device\n"); return 3; }
  stemk_params p;
  memset(&p, 0, sizeof(p));
  p.kind = STEMK_SU_STEM_STR; p.len_band = 10; p.loop_gap = 0.2; p.beta = 0.3; p.stack = 1.3; p.covar = 0.8;
  p.gap = 0.8; p.alpha = 0.2; p.match = 1.0; p.mismatch = 0.8;
  stemk_ctx* ctx[8];
  stemk_set* set[8];
  if (n_dev > 8) n_dev = 8;
  for (int k = 0; k < n_dev; ++k)
    if (stemk_create(&ctx[k], &p, k) != STEMK_OK) { fprintf(stderr, "create %d: %s\n", k, stemk_last_error(NULL)); return 1; }
  if (stemk_upload_multi(ctx, n_dev, &d, set) != STEMK_OK) { fprintf(stderr, "upload_multi: %s\n", stemk_last_error(ctx[0])); return 1; }
  const size_t n = ns;
  double* gm = malloc(8 * n * n), *g1 = malloc(8 * n * n), *g2 = malloc(8 * n * n);
  if (stemk_gram_multi(ctx, (const stemk_set* const*)set, n_dev, 1, gm) != STEMK_OK) { fprintf(stderr, "gram_multi: %s\n", stemk_last_error(ctx[0])); return 1; }
  if (stemk_gram(ctx[0], set[0], 1, g1) != STEMK_OK) { fprintf(stderr, "gram: %s\n", stemk_last_error(ctx[0])); return 1; }
  size_t diff = 0;
  for (size_t k = 0; k < n * n; ++k) diff += memcmp(&gm[k], &g1[k], 8) != 0;
  /* export -> import on the last device -> the same matrix again */
  const int last = n_dev - 1;
  const uint64_t bytes = stemk_set_export_bytes(set[0]);
  void* dbuf = NULL;
  stemk_set* imp = NULL;
  size_t diff2 = 0;
  cudaSetDevice(0);
  if (cudaMalloc(&dbuf, bytes) != 0) { fprintf(stderr, "cudaMalloc failed\n"); return 1; }
  if (stemk_set_export(ctx[0], set[0], dbuf, NULL) != STEMK_OK) { fprintf(stderr, "export: %s\n", stemk_last_error(ctx[0])); return 1; }
  if (last == 0) {
    if (stemk_set_import(ctx[0], dbuf, bytes, &imp) != STEMK_OK) { fprintf(stderr, "import: %s\n", stemk_last_error(ctx[0])); return 1; }
    if (stemk_gram(ctx[0], imp, 1, g2) != STEMK_OK) { fprintf(stderr, "gram(imported): %s\n", stemk_last_error(ctx[0])); return 1; }
    for (size_t k = 0; k < n * n; ++k) diff2 += memcmp(&g2[k], &g1[k], 8) != 0;
    stemk_set_free(ctx[0], imp);
  }
  cudaFree(dbuf);
  printf("%s devices %d records %zu entries_differing %zu import_differing %zu export_bytes %llu\n",
         diff == 0 && diff2 == 0 ? "GRAM_MULTI OK" : "GRAM_MULTI FAILED", n_dev, n, diff, diff2, (unsigned long long)bytes);
  for (int k = 0; k < n_dev; ++k) { stemk_set_free(ctx[k], set[k]); stemk_destroy(ctx[k]); }
  return diff == 0 && diff2 == 0 ? 0 : 1;
}
