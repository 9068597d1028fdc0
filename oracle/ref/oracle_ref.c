/* oracle_ref.c -- TEST INFRASTRUCTURE: literal C restatement of forge-ec's OWN secp256k1 code path.
 *
 * Only tests/ and the cpu_baseline / --impl reference legs of bench.py may call this; the product
 * never does.  The reference is Rust and there is no Rust toolchain in this image, so its real CPU
 * path cannot be executed here; this file follows it line by line (release-profile wrapping
 * semantics) so that (a) the "reference on CPU" timing has the reference's own operation counts
 * (256 ladder steps x {add, 2 doublings}, one Fermat inversion per point) and (b) the divergence
 * between what forge-ec computes and what it specifies can be reported.
 * It is NOT the oracle the CUDA path is graded against (that is oracle/std): the reference's limb
 * code does not implement secp256k1 (mul drops a carry, square is not mul(a,a), R^2 is wrong, the
 * ladder reads a byte-reversed scalar).
 *
 * Every function cites the lines it restates in forge-ec-curves/src/secp256k1.rs.
 * Pinned by tests/test_oracle.py against the values the reference's own tests hold
 * (secp256k1.rs:2750-2760: (1*2).limb0 = 12713681792961361445, 2.square() = 4) and the
 * transliteration goldens of SURVEY.md Appendix A.2 ("parity unpinned against rustc": no
 * toolchain here can run the Rust itself).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint64_t u64;
typedef unsigned __int128 u128;
typedef struct { u64 l[4]; } fe;

static const u64 P[4] = {0xFFFFFFFEFFFFFC2FULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL}; /* :22-23 */
static const u64 N[4] = {0xBFD25E8CD0364141ULL, 0xBAAEDCE6AF48A03BULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFEULL}; /* :27-28 */
static const u64 N0 = 0xD838091DD2253531ULL;                                                                       /* :270, :473 */

/* compare_with_p  :47-75 */
static int compare_with_p(const u64* a) {
  int result = 0;
  for (int i = 3; i >= 0; i--) {
    int nr = a[i] < P[i] ? -1 : (a[i] > P[i] ? 1 : 0);
    if (result == 0) result = nr;
  }
  return result;
}
/* reduce  :78-102 : one conditional subtraction of p */
static void fe_reduce(fe* a) {
  fe r = *a;
  u64 borrow = 0;
  for (int i = 0; i < 4; i++) {
    u64 d1 = r.l[i] - P[i];  int b1 = r.l[i] < P[i];
    u64 d2 = d1 - borrow;    int b2 = d1 < borrow;
    r.l[i] = d2;
    borrow = (u64)(b1 | b2);
  }
  if (compare_with_p(a->l) >= 0) *a = r;
}
static int fe_is_zero(const fe* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }   /* :595-597 */
static int fe_eq(const fe* a, const fe* b) { return memcmp(a->l, b->l, 32) == 0; }            /* :344-351 */

/* impl Add  :353-393 */
static fe fe_add(fe a, fe b) {
  fe result = a;
  u64 carry = 0;
  for (int i = 0; i < 4; i++) {
    u64 sum1 = result.l[i] + b.l[i];
    u64 sum2 = sum1 + carry;
    u64 carry1 = result.l[i] > ~b.l[i];
    u64 carry2 = sum1 > (~(u64)0 - carry);
    result.l[i] = sum2;
    carry = carry1 | carry2;
  }
  fe reduced = result;
  u64 borrow = 0;
  for (int i = 0; i < 4; i++) {
    u64 d1 = reduced.l[i] - P[i];  int b1 = reduced.l[i] < P[i];
    u64 d2 = d1 - borrow;          int b2 = d1 < borrow;
    reduced.l[i] = d2;
    borrow = (u64)(b1 | b2);
  }
  return (carry > 0 || compare_with_p(result.l) >= 0) ? reduced : result;
}
/* impl Sub  :395-440 */
static fe fe_sub(fe a, fe b) {
  fe result = a;
  u64 borrow = 0;
  for (int i = 0; i < 4; i++) {
    u64 d1 = result.l[i] - b.l[i];
    u64 d2 = d1 - borrow;
    u64 b1 = result.l[i] < b.l[i];
    u64 b2 = d1 < borrow;
    result.l[i] = d2;
    borrow = b1 | b2;
  }
  fe wp = result;
  u64 carry = 0;
  for (int i = 0; i < 4; i++) {
    u64 s1 = wp.l[i] + P[i];
    u64 s2 = s1 + carry;
    u64 c1 = wp.l[i] > ~P[i];
    u64 c2 = s1 > (~(u64)0 - carry);
    wp.l[i] = s2;
    carry = c1 | c2;
  }
  return borrow > 0 ? wp : result;
}
/* the shared reduction loop of Mul (:470-494) and mont_reduce_internal (:270-294):
 * `carry` survives from one outer iteration into the next and whatever is left after t[7] is lost */
static fe mont_tail(u64 t[8]) {
  u64 carry = 0;
  for (int i = 0; i < 4; i++) {
    u64 m = t[i] * N0;
    u128 sum = (u128)t[i] + (u128)m * P[0] + carry;
    carry = (u64)(sum >> 64);
    for (int j = 1; j < 4; j++) {
      sum = (u128)t[i + j] + (u128)m * P[j] + carry;
      t[i + j] = (u64)sum;
      carry = (u64)(sum >> 64);
    }
    int j = i + 4;
    while (j < 8 && carry > 0) {
      u128 s = (u128)t[j] + carry;
      t[j] = (u64)s;
      carry = (u64)(s >> 64);
      j++;
    }
  }
  fe r = {{t[4], t[5], t[6], t[7]}};
  if (compare_with_p(r.l) >= 0) fe_reduce(&r);
  return r;
}
/* impl Mul  :442-507 */
static fe fe_mul(fe a, fe b) {
  u64 t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u64 carry = 0;
    for (int j = 0; j < 4; j++) {
      u128 product = (u128)a.l[i] * b.l[j] + t[i + j] + carry;
      t[i + j] = (u64)product;
      carry = (u64)(product >> 64);
    }
    t[i + 4] = carry;
  }
  return mont_tail(t);
}
/* mont_reduce  :314-331 (used by to_bytes :138-179) */
static fe fe_mont_reduce(fe a) {
  u64 t[8] = {a.l[0], a.l[1], a.l[2], a.l[3], 0, 0, 0, 0};
  return mont_tail(t);
}
/* to_montgomery  :219-235 with the reference's R_SQUARED constant */
static fe fe_to_montgomery(fe a) {
  const fe r2 = {{0x000E9F61ULL, 0x07A20000ULL, 0x00000100ULL, 0x00000000ULL}};
  return fe_mul(a, r2);
}
/* square  :634-713 : NOT a Montgomery product; doubled cross terms truncated to 128 bits, the fold
 * multiplies by 0x1000003D1 with 64-bit wrap-around */
static fe fe_square(fe a) {
  u64 product[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 sq = (u128)a.l[i] * a.l[i];
    product[2 * i] = (u64)sq;
    product[2 * i + 1] = (u64)(sq >> 64);
  }
  for (int i = 0; i < 4; i++)
    for (int j = i + 1; j < 4; j++) {
      u128 cross = (u128)a.l[i] * a.l[j];
      cross = cross * 2;                                  /* wrapping_mul(2) on u128 */
      u64 lo = (u64)cross, hi = (u64)(cross >> 64);
      u64 s = product[i + j] + lo;   int carry = s < lo;
      product[i + j] = s;
      u64 s2 = product[i + j + 1] + hi;  int carry2 = s2 < hi;
      product[i + j + 1] = s2;
      if (carry || carry2) {
        int k = i + j + 2;
        while (k < 8) {
          product[k] = product[k] + 1;
          if (product[k] != 0) break;
          k++;
        }
      }
    }
  u64 result[4] = {product[0], product[1], product[2], product[3]};
  u64 carry = 0;
  for (int i = 4; i < 8; i++) {
    u64 w = product[i] * 0x1000003D1ULL;
    u64 t = result[0] + w;
    t = t + carry;
    result[0] = t;
    carry = (u64)(t < w) | ((u64)(t < carry) & (u64)(w != 0));
    for (int j = 1; j < 4; j++) {
      u64 tt = result[j] + carry;
      result[j] = tt;
      carry = tt < carry;
    }
  }
  fe r = {{result[0], result[1], result[2], result[3]}};
  fe_reduce(&r);
  return r;
}
static fe fe_one(void) { fe r = {{1, 0, 0, 0}}; return r; }       /* one()  :585-593 */
static fe fe_zero(void) { fe r = {{0, 0, 0, 0}}; return r; }
/* invert  :599-632 : limbs visited in the order 0,1,2,3, each from bit 63 down */
static fe fe_invert(fe a) {
  static const u64 e[4] = {0xFFFFFFFEFFFFFC2DULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL};
  fe result = fe_one();
  for (int i = 0; i < 4; i++)
    for (int j = 63; j >= 0; j--) {
      result = fe_square(result);
      if ((e[i] >> j) & 1) result = fe_mul(result, a);
    }
  return result;
}
/* FieldElement::from_bytes  :182-213 (big-endian, to_montgomery; no range rejection needed here) */
static fe fe_from_bytes(const uint8_t* b) {
  fe r = fe_zero();
  for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) r.l[3 - i] |= (u64)b[i * 8 + j] << (56 - j * 8);
  return fe_to_montgomery(r);
}
/* FieldElement::to_bytes  :138-179 */
static void fe_to_bytes(uint8_t* out, fe a) {
  fe t = fe_mont_reduce(a);
  for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) out[i * 8 + j] = (uint8_t)(t.l[3 - i] >> (56 - j * 8));
}

typedef struct { fe x, y, z; } pt;
static pt pt_identity(void) { pt r = {fe_zero(), fe_one(), fe_zero()}; return r; }   /* :1322-1324 */
static int pt_is_identity(const pt* p) { return fe_is_zero(&p->z); }                 /* :1326-1340 (both branches reduce to z == 0) */
/* inherent double  :1502-1540 */
static pt pt_double(const pt* p) {
  if (pt_is_identity(p)) return pt_identity();
  fe a = fe_square(p->x), b = fe_square(p->y), c = fe_square(b);
  fe xb = fe_add(p->x, b), xb2 = fe_square(xb);
  fe d = fe_sub(fe_sub(xb2, a), c); d = fe_add(d, d);
  fe three = {{3, 0, 0, 0}}, eight = {{8, 0, 0, 0}};
  fe e = fe_mul(a, three), f = fe_square(e);
  pt r;
  r.x = fe_sub(f, fe_add(d, d));
  r.y = fe_sub(fe_mul(e, fe_sub(d, r.x)), fe_mul(c, eight));
  fe yz = fe_mul(p->y, p->z);
  r.z = fe_add(yz, yz);
  return r;
}
/* impl Add  :1447-1497 */
static pt pt_add(const pt* p, const pt* q) {
  if (pt_is_identity(p)) return *q;
  if (pt_is_identity(q)) return *p;
  fe z1s = fe_square(p->z), z2s = fe_square(q->z);
  fe u1 = fe_mul(p->x, z2s), u2 = fe_mul(q->x, z1s);
  fe z1c = fe_mul(z1s, p->z), z2c = fe_mul(z2s, q->z);
  fe s1 = fe_mul(p->y, z2c), s2 = fe_mul(q->y, z1c);
  if (fe_eq(&u1, &u2)) {
    if (fe_eq(&s1, &s2)) return pt_double(p);
    return pt_identity();
  }
  fe h = fe_sub(u2, u1), r = fe_sub(s2, s1);
  fe h2 = fe_square(h), h3 = fe_mul(h2, h), u1h2 = fe_mul(u1, h2);
  pt o;
  o.x = fe_sub(fe_sub(fe_sub(fe_square(r), h3), u1h2), u1h2);
  o.y = fe_sub(fe_mul(r, fe_sub(u1h2, o.x)), fe_mul(s1, h3));
  o.z = fe_mul(fe_mul(h, p->z), q->z);
  return o;
}
/* Secp256k1::generator  :2608-2625 */
static pt generator(void) {
  fe gx = {{0x59F2815B16F81798ULL, 0x029BFCDB2DCE28D9ULL, 0x55A06295CE870B07ULL, 0x79BE667EF9DCBBACULL}};
  fe gy = {{0x9C47D08FFB10D4B8ULL, 0xFD17B448A6855419ULL, 0x5DA4FBFC0E1108A8ULL, 0x483ADA7726A3C465ULL}};
  pt g = {fe_to_montgomery(gx), fe_to_montgomery(gy), fe_one()};
  return g;
}
/* Secp256k1::multiply  :2635-2692.  scalar limbs come from Scalar::from_bytes (BE, :2271-2297);
 * scalar.to_bytes() resolves to the inherent little-endian version (:1924-1933), and the ladder then
 * reads those bytes "MSB first" (:2655-2659) */
static pt multiply(const pt* point, const u64 k[4]) {
  if (pt_is_identity(point) || (k[0] | k[1] | k[2] | k[3]) == 0) return pt_identity();
  uint8_t sb[32];
  for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) sb[i * 8 + j] = (uint8_t)(k[i] >> (j * 8));
  pt r0 = pt_identity(), r1 = *point;
  for (int i = 0; i < 256; i++) {
    int bit = (sb[i / 8] >> (7 - (i % 8))) & 1;
    pt sum = pt_add(&r0, &r1);
    pt d0 = pt_double(&r0), d1 = pt_double(&r1);
    r0 = bit ? sum : d0;       /* conditional_select(a, b, c) returns b iff c == 1 */
    r1 = bit ? d1 : sum;
  }
  return r0;
}
/* to_affine  :1342-1363 + AffinePoint::to_bytes  :875-894 */
static void to_affine_bytes(uint8_t out[33], const pt* p) {
  memset(out, 0, 33);
  if (pt_is_identity(p)) return;
  fe zi = fe_invert(p->z);
  fe zi2 = fe_square(zi), zi3 = fe_mul(zi2, zi);
  fe x = fe_mul(p->x, zi2), y = fe_mul(p->y, zi3);
  uint8_t yb[32];
  fe_to_bytes(yb, y);
  out[0] = (yb[31] & 1) ? 3 : 2;
  fe_to_bytes(out + 1, x);
}

static void lane(const uint8_t* sc, const uint8_t* xy, uint8_t* out) {
  u64 k[4] = {0, 0, 0, 0};
  for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) k[i] |= (u64)sc[31 - (i * 8 + j)] << (j * 8);
  int valid = !(k[3] > N[3] || (k[3] == N[3] && k[2] > N[2]) || (k[3] == N[3] && k[2] == N[2] && k[1] > N[1]) ||
                (k[3] == N[3] && k[2] == N[2] && k[1] == N[1] && k[0] >= N[0]));
  if (!valid) { memset(out, 0, 33); return; }   /* CtOption::none -> the caller has no scalar */
  pt p;
  if (xy == NULL) {
    p = generator();
  } else {
    p.x = fe_from_bytes(xy); p.y = fe_from_bytes(xy + 32); p.z = fe_one();   /* from_affine :1365-1373 */
  }
  pt r = multiply(&p, k);
  to_affine_bytes(out, &r);
}

typedef struct { const uint8_t *sc, *xy; uint8_t* out; size_t first, count; } job_t;
static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  for (size_t i = j->first; i < j->first + j->count; i++) lane(j->sc + 32 * i, j->xy ? j->xy + 64 * i : NULL, j->out + 33 * i);
  return NULL;
}
/* n lanes of: Scalar::from_bytes, from_affine | generator, multiply, to_affine, to_bytes */
void oref_secp256k1_batch_mul(const uint8_t* scalars_be, const uint8_t* points_xy_be, size_t n, uint8_t* out33, int threads) {
  if (threads < 1) threads = 1;
  if ((size_t)threads > n) threads = n ? (int)n : 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
  job_t* jobs = (job_t*)malloc(sizeof(job_t) * (size_t)threads);
  for (int t = 0; t < threads; t++) {
    jobs[t].sc = scalars_be; jobs[t].xy = points_xy_be; jobs[t].out = out33;
    jobs[t].first = n * (size_t)t / (size_t)threads;
    jobs[t].count = n * (size_t)(t + 1) / (size_t)threads - jobs[t].first;
    if (threads == 1) worker(&jobs[t]); else pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  if (threads > 1) for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
  free(th); free(jobs);
}
/* probes for the known-answer tests: op 0 mul, 1 square, 2 add, 3 sub on raw limbs (little-endian u64 x 4) */
void oref_secp256k1_field_op(int op, const u64* a, const u64* b, u64* out) {
  fe x, y, r;
  memcpy(x.l, a, 32); memcpy(y.l, b, 32);
  r = op == 0 ? fe_mul(x, y) : op == 1 ? fe_square(x) : op == 2 ? fe_add(x, y) : fe_sub(x, y);
  memcpy(out, r.l, 32);
}
void oref_secp256k1_generator_x_raw(u64* out) { pt g = generator(); memcpy(out, g.x.l, 32); }
