/* oracle_ref_x25519.c -- TEST INFRASTRUCTURE: literal C restatement of forge-ec's OWN x25519() code path
 * (forge-ec-curves/src/curve25519.rs:1624-1716 and the FieldElement it runs on, :50-114, 117-163, 186-334, 355-494),
 * the companion of oracle_ref.c / oracle_ref_p256.c.
 *
 * Only tests/ may call this; the product never does.  The reference's 2^255-19 code is limb arithmetic without
 * carries between the limbs (Add, Sub, Neg and the final subtraction of p work limb by limb with wrapping), a
 * multiplication that drops carries and folds the high 256 bits with 19 instead of 38, and a big-endian byte codec
 * where RFC 7748 is little-endian; x25519() multiplies by A = 486662 where the ladder needs a24 and returns a
 * hard-coded value for the scalar 2.  Its outputs are therefore not RFC 7748 results; this file
 * reproduces them bit for bit, in the same order of operations and with release-profile wrapping semantics.
 *
 * Pinned by tests/test_oracle.py against SURVEY.md Appendix A.2 (the RFC 7748 Alice scalar, the scalars 1 and 2 on
 * u = 9; values from the survey's independent transliteration -- "not verified against rustc": no toolchain here can
 * run the Rust, and the reference's own tests never compare a computed X25519 output with anything, :2107-2124).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint64_t u64;
typedef unsigned __int128 u128;
typedef struct { u64 l[4]; } fe;

static const u64 P[4] = {0xFFFFFFFFFFFFFFEDULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL, 0x7FFFFFFFFFFFFFFFULL}; /* :20-21 */

static int ge_p(const fe* a) {   /* the comparison of reduce :95-105 and from_bytes :148-156 */
  int ge_p3 = a->l[3] > P[3], eq_p3 = a->l[3] == P[3], eq_p2 = a->l[2] == P[2], eq_p1 = a->l[1] == P[1], ge_p0 = a->l[0] >= P[0];
  return ge_p3 | (eq_p3 & eq_p2 & eq_p1 & ge_p0);
}
/* reduce :50-114 */
static void fe_reduce(fe* a) {
  u64 bit255 = (a->l[3] >> 63) & 1;
  a->l[0] = a->l[0] + bit255 * 19;          /* wrapping, no carry into limb 1 */
  a->l[3] &= 0x7FFFFFFFFFFFFFFFULL;
  u64 carry = 0;
  for (int i = 0; i < 4; i++) {             /* adds a carry that starts at 0: never changes anything (:65-73) */
    u64 sum = a->l[i] + carry;
    u64 c1 = sum < a->l[i];
    a->l[i] = sum;
    carry = c1 ? 1 : 0;
  }
  if (carry > 0) {                          /* unreachable, kept for the line-by-line correspondence (:76-92) */
    u64 sum = a->l[0] + carry * 19;
    u64 c1 = sum < a->l[0];
    a->l[0] = sum;
    if (c1) {
      a->l[1] += 1;
      if (a->l[1] == 0) { a->l[2] += 1; if (a->l[2] == 0) a->l[3] += 1; }
    }
  }
  if (ge_p(a)) for (int i = 0; i < 4; i++) a->l[i] -= P[i];   /* limb by limb, no borrows (:108-113) */
}
static fe fe_from_be(const uint8_t* b, int* valid) {   /* inherent from_bytes :130-163 */
  fe r = {{0, 0, 0, 0}};
  for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) r.l[i] |= (u64)b[31 - (i * 8 + j)] << (j * 8);
  fe_reduce(&r);
  *valid = !ge_p(&r);
  return r;
}
static void fe_to_be(uint8_t* out, fe a) {             /* inherent to_bytes :117-128 (no reduction) */
  for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) out[31 - (i * 8 + j)] = (uint8_t)(a.l[i] >> (j * 8));
}
static fe fe_zero(void) { fe r = {{0, 0, 0, 0}}; return r; }
static fe fe_one(void) { fe r = {{1, 0, 0, 0}}; return r; }
static int fe_is_zero(const fe* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }   /* :363-365 */
/* impl Add :186-203 */
static fe fe_add(fe a, fe b) {
  fe r;
  for (int i = 0; i < 4; i++) r.l[i] = a.l[i] + b.l[i];
  fe_reduce(&r);
  return r;
}
/* impl Sub :205-225 */
static fe fe_sub(fe a, fe b) {
  fe r;
  for (int i = 0; i < 4; i++) r.l[i] = a.l[i] + P[i] - b.l[i];
  fe_reduce(&r);
  return r;
}
/* impl Mul :227-314 */
static fe fe_mul(fe a, fe b) {
  u64 result[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    for (int j = 0; j < 4; j++) {
      u128 product = (u128)a.l[i] * (u128)b.l[j];
      u64 low = (u64)product, high = (u64)(product >> 64);
      int idx = i + j;
      u64 sum1 = result[idx] + low;
      int carry1 = sum1 < result[idx];
      result[idx] = sum1;
      if (carry1) result[idx + 1] = result[idx + 1] + 1;          /* its own overflow is dropped */
      u64 sum3 = result[idx + 1] + high;
      int carry3 = sum3 < result[idx + 1];
      result[idx + 1] = sum3;
      if (carry3 && idx + 2 < 8) result[idx + 2] += 1;             /* `+=`: wraps in release mode */
    }
  }
  for (int i = 0; i < 4; i++) {
    u64 high_bits = result[4 + i];
    if (high_bits > 0) {
      u128 product = (u128)high_bits * 19;
      u64 low = (u64)product, high = (u64)(product >> 64);
      u64 sum1 = result[i] + low;
      int carry1 = sum1 < result[i];
      result[i] = sum1;
      u64 carry = carry1 ? 1 : 0;
      if (high > 0) {
        u64 sum2 = result[i + 1] + high;
        int carry2 = sum2 < result[i + 1];
        result[i + 1] = sum2;
        if (carry2 && i + 2 < 4) result[i + 2] += 1;
      }
      if (carry > 0) {
        u64 sum3 = result[i + 1] + carry;
        int carry3 = sum3 < result[i + 1];
        result[i + 1] = sum3;
        if (carry3 && i + 2 < 4) result[i + 2] += 1;
      }
    }
  }
  fe r = {{result[0], result[1], result[2], result[3]}};
  fe_reduce(&r);
  return r;
}
static fe fe_square(fe a) { return fe_mul(a, a); }   /* :490-494 */
static fe fe_sqn(fe a, int n) { for (int i = 0; i < n; i++) a = fe_square(a); return a; }
/* invert :367-488: the reference's hand-made chain, whatever exponent it amounts to */
static fe fe_invert(fe a, int* some) {
  if (fe_is_zero(&a)) { *some = 0; return fe_zero(); }
  *some = 1;
  fe a2 = fe_square(a), a4 = fe_square(a2), a16 = fe_sqn(a4, 2);
  fe a256 = fe_sqn(a16, 4);        /* one squaring, then three more (:399-402) */
  fe a65536 = fe_sqn(a256, 8);
  fe t32 = fe_sqn(a65536, 16), t64 = fe_sqn(t32, 32), t128 = fe_sqn(t64, 64), t192 = fe_sqn(t128, 64), t250 = fe_sqn(t192, 58);
  fe result = fe_mul(t250, a);
  result = fe_mul(result, a2);
  result = fe_mul(result, a4);
  result = fe_mul(result, fe_square(a4));
  result = fe_mul(result, a16);
  for (int i = 0; i < 4; i++) result = fe_mul(result, result);
  fe a64 = fe_sqn(a16, 2), a32 = fe_square(a16), a8 = fe_square(a4);
  result = fe_mul(fe_mul(fe_mul(fe_mul(fe_mul(result, a64), a32), a8), a2), a);
  return result;
}
/* x25519 :1624-1716 */
static void x25519_ref(const uint8_t scalar[32], const uint8_t u[32], uint8_t out[32]) {
  int two = scalar[0] == 2;
  for (int i = 1; i < 32; i++) two &= scalar[i] == 0;
  if (two) {
    static const uint8_t fixed[32] = {0x1b, 0x7f, 0x9f, 0x7c, 0x27, 0x65, 0x50, 0xbb, 0x3a, 0x3c, 0xec, 0xc8, 0xa5, 0x77, 0x0c, 0x17,
                                      0x3f, 0x58, 0x31, 0xed, 0x1b, 0xb2, 0x8c, 0x05, 0x58, 0xaa, 0xc4, 0x71, 0x3f, 0x97, 0x08, 0x22};
    memcpy(out, fixed, 32);
    return;
  }
  uint8_t sb[32], ub[32];
  memcpy(sb, scalar, 32);
  sb[0] &= 248; sb[31] &= 127; sb[31] |= 64;
  memcpy(ub, u, 32);
  ub[31] &= 127;
  int valid;
  fe u_fe = fe_from_be(ub, &valid);          /* the BIG-endian inherent from_bytes (:132) shadows the trait's */
  if (!valid) u_fe = fe_zero();
  fe x1 = u_fe, x2 = fe_one(), z2 = fe_zero(), x3 = u_fe, z3 = fe_one();
  const fe A = {{0x0000000000076D06ULL, 0, 0, 0}};
  unsigned swap = 0;
  for (int i = 254; i >= 0; i--) {
    unsigned bit = (sb[i / 8] >> (i % 8)) & 1;
    unsigned new_swap = swap ^ bit;
    if (new_swap) { fe tx = x2, tz = z2; x2 = x3; z2 = z3; x3 = tx; z3 = tz; }
    swap = bit;
    fe a = fe_add(x2, z2), aa = fe_square(a), b = fe_sub(x2, z2), bb = fe_square(b), e = fe_sub(aa, bb);
    fe c = fe_add(x3, z3), d = fe_sub(x3, z3), da = fe_mul(d, a), cb = fe_mul(c, b);
    x3 = fe_square(fe_add(da, cb));
    z3 = fe_mul(x1, fe_square(fe_sub(da, cb)));
    x2 = fe_mul(aa, bb);
    z2 = fe_mul(e, fe_add(aa, fe_mul(A, e)));
  }
  if (swap) { x2 = x3; z2 = z3; }
  int some;
  fe z2_inv = fe_invert(z2, &some);
  if (!some) z2_inv = fe_zero();
  fe_to_be(out, fe_mul(x2, z2_inv));          /* the BIG-endian inherent to_bytes (:117), no final reduction */
}

typedef struct { const uint8_t *k, *u; uint8_t* out; size_t first, count; } job_t;
static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  static const uint8_t nine[32] = {9};
  for (size_t i = j->first; i < j->first + j->count; i++) x25519_ref(j->k + 32 * i, j->u ? j->u + 32 * i : nine, j->out + 32 * i);
  return NULL;
}
/* n lanes of x25519(scalar_i, u_i); u == NULL: the base point [9, 0, ..] (Curve25519::generator, :1900-1904) */
void oref_x25519_batch(const uint8_t* scalars, const uint8_t* us, size_t n, uint8_t* out, int threads) {
  if (threads < 1) threads = 1;
  if ((size_t)threads > n) threads = n ? (int)n : 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
  job_t* jobs = (job_t*)malloc(sizeof(job_t) * (size_t)threads);
  for (int t = 0; t < threads; t++) {
    jobs[t].k = scalars; jobs[t].u = us; jobs[t].out = out;
    jobs[t].first = n * (size_t)t / (size_t)threads;
    jobs[t].count = n * (size_t)(t + 1) / (size_t)threads - jobs[t].first;
    if (threads == 1) worker(&jobs[t]); else pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  if (threads > 1) for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
  free(th); free(jobs);
}
/* probes for the tests: op 0 mul, 1 square, 2 add, 3 sub on raw limbs (little-endian u64 x 4) */
void oref_25519_field_op(int op, const u64* a, const u64* b, u64* out) {
  fe x, y, r;
  memcpy(x.l, a, 32); memcpy(y.l, b, 32);
  r = op == 0 ? fe_mul(x, y) : op == 1 ? fe_square(x) : op == 2 ? fe_add(x, y) : fe_sub(x, y);
  memcpy(out, r.l, 32);
}
