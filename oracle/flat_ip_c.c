/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/flat_ip.py for the parity status: the FAISS
 * arithmetic is unpinned because faiss is absent from /root/reference and from this image).
 *
 * C restatement of the scalar branch of FAISS 1.7.x `knn_inner_product` (the `nq < 20` path of
 * utils/distances.cpp that `faiss.IndexFlatIP.search` takes; reference call site
 * drivers/run_convdr_inference.py:182): for every query, one pass over the database computing the
 * fp32 inner product, a k-entry MIN-heap whose root is replaced only on a strictly larger score,
 * and a final reorder to descending.  OpenMP over queries, like upstream.
 *
 * Also: the host twin of the device synthetic-row generator (convdr_b200/csrc/kernels_util.cuh,
 * convdr_b200/synth.py), used to materialise CPU-baseline samples quickly.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static float dot_f32(const float* a, const float* b, int d) {
  /* 8 independent partial sums, like an 8-lane SIMD accumulator, then a horizontal add */
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int t = 0;
  for (; t + 8 <= d; t += 8)
    for (int l = 0; l < 8; ++l) acc[l] += a[t + l] * b[t + l];
  float s = ((acc[0] + acc[4]) + (acc[1] + acc[5])) + ((acc[2] + acc[6]) + (acc[3] + acc[7]));
  for (; t < d; ++t) s += a[t] * b[t];
  return s;
}

/* min-heap on (val, id), 0-based; root = current k-th best */
static void sift_down(float* val, int64_t* ids, int k, int i) {
  const float v = val[i];
  const int64_t id = ids[i];
  for (;;) {
    int c = 2 * i + 1;
    if (c >= k) break;
    if (c + 1 < k && val[c + 1] < val[c]) c += 1;
    if (!(val[c] < v)) break;
    val[i] = val[c];
    ids[i] = ids[c];
    i = c;
  }
  val[i] = v;
  ids[i] = id;
}

void oracle_knn_ip_heap(const float* x, const float* xb, int d, int64_t nq, int64_t n, int k, float* D,
                        int64_t* I) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t q = 0; q < nq; ++q) {
    float* val = D + q * k;
    int64_t* ids = I + q * k;
    for (int i = 0; i < k; ++i) { val[i] = -FLT_MAX; ids[i] = -1; }
    const float* xq = x + q * d;
    for (int64_t j = 0; j < n; ++j) {
      const float ip = dot_f32(xq, xb + j * d, d);
      if (ip > val[0]) {            /* strict: at equal score the earlier (lower) index stays */
        val[0] = ip;
        ids[0] = j;
        sift_down(val, ids, k, 0);
      }
    }
    /* reorder: pop the minimum to the back until the array is descending */
    for (int m = k; m > 1; --m) {
      float tv = val[0]; int64_t ti = ids[0];
      val[0] = val[m - 1]; ids[0] = ids[m - 1];
      val[m - 1] = tv; ids[m - 1] = ti;
      sift_down(val, ids, m - 1, 0);
    }
  }
}

/* ---- synthetic rows: Philox4x32-10 + integer Irwin-Hall + exact normalisation ---- */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

void oracle_synth_rows(float* out, int64_t first_row, int64_t n, uint64_t seed, uint64_t stream, float norm) {
  const uint32_t k0 = (uint32_t)seed ^ (uint32_t)stream;
  const uint32_t k1 = (uint32_t)(seed >> 32) ^ (uint32_t)(stream >> 32) ^ 0x5eedu;
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < n; ++r) {
    const uint64_t row = (uint64_t)(first_row + r);
    int comp[768];
    int64_t ss = 0;
    for (uint32_t ch = 0; ch < 192; ++ch) {
      uint32_t c[4] = {(uint32_t)row, (uint32_t)(row >> 32), ch, 0u};
      philox4x32_10(c, k0, k1);
      for (int j = 0; j < 4; ++j) {
        const uint32_t w = c[j];
        const int v = (int)((w & 0xffu) + ((w >> 8) & 0xffu) + ((w >> 16) & 0xffu) + (w >> 24)) - 510;
        comp[4 * ch + j] = v;
        ss += (int64_t)v * v;
      }
    }
    volatile float root = sqrtf((float)ss);   /* IEEE correctly rounded; volatile blocks rsqrt tricks */
    const float inv = ss > 0 ? norm / root : 0.0f;
    float* o = out + r * 768;
    for (int t = 0; t < 768; ++t) o[t] = (float)comp[t] * inv;
  }
}
