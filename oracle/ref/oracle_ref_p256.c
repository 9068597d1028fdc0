/* oracle_ref_p256.c -- TEST INFRASTRUCTURE: literal C restatement of forge-ec's OWN P-256 code path
 * (forge-ec-curves/src/p256.rs), the companion of oracle_ref.c (secp256k1).
 *
 * Only tests/ may call this; the product never does.  The reference is Rust and cannot be run here, so this file
 * follows it line by line with release-profile wrapping semantics.  What the reference's P-256 code really does
 * (SURVEY.md section 0): the field is almost right -- Add, Mul (schoolbook + FIPS 186-4 Solinas sums) and the
 * LSB-first pow are correct up to corner cases of probability ~2^-30 per operation (the final +-p loops of
 * reduce_wide_p256 work modulo 2^256) -- but Sub adds p through `+=`, which reduces it away again, so a - b with a < b
 * yields a - b + 2^256 instead of a - b + p, and `double` uses the a = 0 formula (3 X^2 without the a Z^4 term) on a
 * curve with a = -3.  Its outputs are therefore not FIPS 186-4 results; this file reproduces them bit for bit.
 *
 * Every function cites the lines it restates.  Pinned by tests/test_oracle.py against SURVEY.md Appendix A.2
 * (k = 1, 2, 3, 0xdeadbeef; values from the survey's independent transliteration -- "not verified against rustc",
 * no toolchain here can run the Rust) and against a big-integer model of the same semantics (oracle/pymodel.py).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint64_t u64;
typedef unsigned __int128 u128;
typedef __int128 i128;
typedef struct { u64 l[4]; } fe;
typedef struct { fe x, y, z; } pt;

static const u64 P[4] = {0xFFFFFFFFFFFFFFFFULL, 0x00000000FFFFFFFFULL, 0x0000000000000000ULL, 0xFFFFFFFF00000001ULL}; /* p256.rs:18-19 */
static const u64 N[4] = {0xF3B9CAC2FC632551ULL, 0xBCE6FAADA7179E84ULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFF00000000ULL}; /* :23-24 */

/* compare :70-80 */
static int cmp4(const u64* a, const u64* b) {
  for (int i = 3; i >= 0; i--) {
    if (a[i] < b[i]) return -1;
    if (a[i] > b[i]) return 1;
  }
  return 0;
}
/* a - p, wrapping (the subtraction loop of reduce :88-99, Add :458-465, reduce_wide_p256 :672-681) */
static void sub_p(fe* a) {
  u64 borrow = 0;
  for (int i = 0; i < 4; i++) {
    u64 d1 = a->l[i] - P[i], b1 = a->l[i] < P[i];
    u64 d2 = d1 - borrow, b2 = d1 < borrow;
    a->l[i] = d2;
    borrow = b1 + b2;
  }
}
/* reduce :88-100 */
static void fe_reduce(fe* a) { while (cmp4(a->l, P) >= 0) sub_p(a); }
static int fe_is_zero(const fe* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }   /* :763-765 */
static int fe_eq(const fe* a, const fe* b) { return memcmp(a->l, b->l, 32) == 0; }            /* ct_eq :407-414 */
static fe fe_u64(u64 v) { fe r = {{v, 0, 0, 0}}; return r; }                                   /* From<u64> :869-873 */
static int fe_is_one(const fe* a) { fe o = fe_u64(1); return fe_eq(a, &o); }                  /* :1342-1344 */

/* impl Add :416-468 */
static fe fe_add(fe a, fe b) {
  fe r = a;
  u64 carry = 0;
  for (int i = 0; i < 4; i++) {
    u64 s1 = r.l[i] + b.l[i], o1 = s1 < r.l[i];
    u64 s2 = s1 + carry, o2 = s2 < s1;
    r.l[i] = s2;
    carry = o1 + o2;
  }
  while (carry > 0) {
    static const u64 red[4] = {0x0000000000000001ULL, 0xFFFFFFFF00000000ULL, 0xFFFFFFFFFFFFFFFFULL, 0x00000000FFFFFFFEULL};
    u64 add_carry = 0;
    for (int i = 0; i < 4; i++) {
      u64 s1 = r.l[i] + red[i], o1 = s1 < r.l[i];
      u64 s2 = s1 + add_carry, o2 = s2 < s1;
      r.l[i] = s2;
      add_carry = o1 + o2;
    }
    carry = carry - 1 + add_carry;
  }
  while (cmp4(r.l, P) >= 0) sub_p(&r);
  return r;
}
/* impl Sub :470-496: `result += P` goes through Add, which reduces the sum again */
static fe fe_sub(fe a, fe b) {
  fe r = a;
  if (cmp4(a.l, b.l) < 0) { fe pp; memcpy(pp.l, P, 32); r = fe_add(r, pp); }
  fe d;
  u64 borrow = 0;
  for (int i = 0; i < 4; i++) {
    u64 d1 = r.l[i] - b.l[i], b1 = r.l[i] < b.l[i];
    u64 d2 = d1 - borrow, b2 = d1 < borrow;
    d.l[i] = d2;
    borrow = b1 + b2;
  }
  return d;
}
/* impl Neg :707-729 */
static fe fe_neg(fe a) {
  if (fe_is_zero(&a)) return a;
  fe r;
  u64 borrow = 0;
  for (int i = 0; i < 4; i++) {
    u64 d1 = P[i] - a.l[i], b1 = P[i] < a.l[i];
    u64 d2 = d1 - borrow, b2 = d1 < borrow;
    r.l[i] = d2;
    borrow = b1 + b2;
  }
  return r;
}
/* reduce_wide_p256 :544-704 */
static fe reduce_wide_p256(const u64 wide[8]) {
  i128 c[16], acc[8];
  for (int i = 0; i < 8; i++) { c[2 * i] = (i128)(wide[i] & 0xFFFFFFFFULL); c[2 * i + 1] = (i128)(wide[i] >> 32); }
  for (int i = 0; i < 8; i++) acc[i] = c[i];                                                        /* s1 */
  acc[3] += 2 * c[11]; acc[4] += 2 * c[12]; acc[5] += 2 * c[13]; acc[6] += 2 * c[14]; acc[7] += 2 * c[15];   /* 2 s2 */
  acc[3] += 2 * c[12]; acc[4] += 2 * c[13]; acc[5] += 2 * c[14]; acc[6] += 2 * c[15];                        /* 2 s3 */
  acc[0] += c[8]; acc[1] += c[9]; acc[2] += c[10]; acc[6] += c[14]; acc[7] += c[15];                         /* s4 */
  acc[0] += c[9]; acc[1] += c[10]; acc[2] += c[11]; acc[3] += c[13]; acc[4] += c[14]; acc[5] += c[15]; acc[6] += c[13]; acc[7] += c[8];  /* s5 */
  acc[0] -= c[11]; acc[1] -= c[12]; acc[2] -= c[13]; acc[6] -= c[8]; acc[7] -= c[10];                        /* s6 */
  acc[0] -= c[12]; acc[1] -= c[13]; acc[2] -= c[14]; acc[3] -= c[15]; acc[6] -= c[9]; acc[7] -= c[11];       /* s7 */
  acc[0] -= c[13]; acc[1] -= c[14]; acc[2] -= c[15]; acc[3] -= c[8]; acc[4] -= c[9]; acc[5] -= c[10]; acc[7] -= c[12];   /* s8 */
  acc[0] -= c[14]; acc[1] -= c[15]; acc[3] -= c[9]; acc[4] -= c[10]; acc[5] -= c[11]; acc[7] -= c[13];       /* s9 */
  for (int i = 0; i < 7; i++) {
    i128 carry = acc[i] >> 32;  /* arithmetic shift, as Rust's i128 >> */
    acc[i] &= 0xFFFFFFFF;
    acc[i + 1] += carry;
  }
  i128 carry = acc[7] >> 32;
  acc[7] &= 0xFFFFFFFF;
  fe r;
  for (int i = 0; i < 4; i++) r.l[i] = (u64)acc[2 * i] | ((u64)acc[2 * i + 1] << 32);
  while (carry > 0) { sub_p(&r); carry -= 1; }
  while (carry < 0) {
    u64 cc = 0;
    for (int i = 0; i < 4; i++) {
      u64 s1 = r.l[i] + P[i], o1 = s1 < r.l[i];
      u64 s2 = s1 + cc, o2 = s2 < s1;
      r.l[i] = s2;
      cc = o1 + o2;
    }
    carry += 1;
  }
  fe_reduce(&r);
  return r;
}
/* impl Mul :498-534 */
static fe fe_mul(fe a, fe b) {
  u64 wide[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 carry = 0;
    for (int j = 0; j < 4; j++) {
      u128 product = (u128)a.l[i] * (u128)b.l[j] + (u128)wide[i + j] + carry;
      wide[i + j] = (u64)product;
      carry = product >> 64;
    }
    u128 sum = (u128)wide[i + 4] + carry;
    wide[i + 4] = (u64)sum;
    if ((sum >> 64) != 0) {
      for (int k = i + 5; k < 8; k++) {
        wide[k] += 1;
        if (wide[k] != 0) break;
      }
    }
  }
  return reduce_wide_p256(wide);
}
static fe fe_square(fe a) { return fe_mul(a, a); }   /* :772-776 */
/* pow :376-393 (LSB first) and invert :343-370 with exp = P - 2 */
static fe fe_pow(fe a, const u64 exp[4]) {
  fe result = fe_u64(1), base = a;
  for (int w = 0; w < 4; w++) {
    u64 e = exp[w];
    for (int i = 0; i < 64; i++) {
      if (e & 1) result = fe_mul(result, base);
      base = fe_square(base);
      e >>= 1;
    }
  }
  return result;
}
static fe fe_invert(fe a) {
  if (fe_is_zero(&a)) return fe_u64(0);
  u64 exp[4] = {P[0] - 2, P[1], P[2], P[3]};  /* P[0] >= 2: the borrow loop stops at the first limb */
  return fe_pow(a, exp);
}
static fe fe_from_bytes(const uint8_t* b) {   /* :303-317 (the < p check is the caller's) */
  fe r = {{0, 0, 0, 0}};
  for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) r.l[i] |= (u64)b[31 - (i * 8 + j)] << (j * 8);
  return r;
}
static void fe_to_bytes(uint8_t* out, fe a) {  /* :285-298 */
  for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) out[31 - (i * 8 + j)] = (uint8_t)(a.l[i] >> (j * 8));
}

static pt pt_identity(void) { pt r = {fe_u64(0), fe_u64(1), fe_u64(0)}; return r; }   /* :1826-1828 */
static int pt_is_identity(const pt* p) { return fe_is_zero(&p->z); }                  /* :1830-1832 */
/* double :1869-1912 -- the a = 0 formula */
static pt pt_double(const pt* p) {
  if (pt_is_identity(p)) return pt_identity();
  fe xx = fe_square(p->x), yy = fe_square(p->y), yyyy = fe_square(yy);
  fe xy2 = fe_square(fe_add(p->x, yy));
  fe t = fe_sub(fe_sub(xy2, xx), yyyy);
  fe d = fe_add(t, t);
  fe e = fe_mul(fe_u64(3), xx);
  fe ee = fe_square(e);
  fe x3 = fe_sub(fe_sub(ee, d), d);
  fe eight_yyyy = fe_mul(fe_u64(8), yyyy);
  fe y3 = fe_sub(fe_mul(e, fe_sub(d, x3)), eight_yyyy);
  fe z3 = fe_add(p->y, p->y);
  if (!fe_is_one(&p->z)) z3 = fe_mul(z3, p->z);
  pt r = {x3, y3, z3};
  return r;
}
/* ConstantTimeEq for ProjectivePoint :2035-2067 */
static int pt_eq(const pt* a, const pt* b) {
  if (pt_is_identity(a) && pt_is_identity(b)) return 1;
  if (pt_is_identity(a) || pt_is_identity(b)) return 0;
  fe z1z1 = fe_square(a->z), z2z2 = fe_square(b->z);
  fe u1 = fe_mul(a->x, z2z2), u2 = fe_mul(b->x, z1z1);
  fe s1 = fe_mul(fe_mul(a->y, b->z), z2z2), s2 = fe_mul(fe_mul(b->y, a->z), z1z1);
  return fe_eq(&u1, &u2) & fe_eq(&s1, &s2);
}
/* impl Add for ProjectivePoint :1941-2006 */
static pt pt_add(const pt* a, const pt* b) {
  if (pt_is_identity(a)) return *b;
  if (pt_is_identity(b)) return *a;
  if (pt_eq(a, b)) return pt_double(a);
  fe z1z1 = fe_square(a->z), z2z2 = fe_square(b->z);
  fe u1 = fe_mul(a->x, z2z2), u2 = fe_mul(b->x, z1z1);
  fe s1 = fe_mul(fe_mul(a->y, b->z), z2z2), s2 = fe_mul(fe_mul(b->y, a->z), z1z1);
  fe ns2 = fe_neg(s2);
  if (fe_eq(&u1, &u2) && fe_eq(&s1, &ns2)) return pt_identity();
  fe h = fe_sub(u2, u1);
  fe i = fe_square(fe_add(h, h));
  fe j = fe_mul(h, i);
  fe r = fe_add(fe_sub(s2, s1), fe_sub(s2, s1));
  fe v = fe_mul(u1, i);
  fe x3 = fe_sub(fe_sub(fe_sub(fe_square(r), j), v), v);
  fe y3 = fe_sub(fe_mul(r, fe_sub(v, x3)), fe_mul(fe_add(s1, s1), j));
  fe z3 = fe_mul(fe_sub(fe_sub(fe_square(fe_add(a->z, b->z)), z1z1), z2z2), h);
  pt o = {x3, y3, z3};
  return o;
}
/* P256::generator :2092-2110 */
static pt generator(void) {
  pt g = {{{0xF4A13945D898C296ULL, 0x77037D812DEB33A0ULL, 0xF8BCE6E563A440F2ULL, 0x6B17D1F2E12C4247ULL}},
          {{0xCBB6406837BF51F5ULL, 0x2BCE33576B315ECEULL, 0x8EE7EB4A7C0F9E16ULL, 0x4FE342E2FE1A7F9BULL}},
          {{1, 0, 0, 0}}};
  return g;
}
/* P256::multiply :2120-2156 -- MSB-first double-and-add over the big-endian bytes of the scalar */
static pt multiply(const pt* point, const uint8_t k_be[32]) {
  int zero = 1;
  for (int i = 0; i < 32; i++) zero &= k_be[i] == 0;
  if (pt_is_identity(point) || zero) return pt_identity();
  pt result = pt_identity();
  for (int i = 0; i < 256; i++) {
    int bit = (k_be[i / 8] >> (7 - (i % 8))) & 1;
    result = pt_double(&result);
    if (bit) result = pt_add(&result, point);
  }
  return result;
}
/* to_affine :1835-1857 + AffinePoint::to_bytes :1558-1577 */
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
  fe k = fe_from_bytes(sc);                        /* Scalar::from_bytes :1041-1055: same byte order, check against n */
  if (cmp4(k.l, N) >= 0) { memset(out, 0, 33); return; }   /* CtOption::none -> the caller has no scalar */
  pt p;
  if (xy == NULL) {
    p = generator();
  } else {
    p.x = fe_from_bytes(xy); p.y = fe_from_bytes(xy + 32); p.z = fe_u64(1);   /* from_affine :1859-1867 */
  }
  pt r = multiply(&p, sc);
  to_affine_bytes(out, &r);
}

typedef struct { const uint8_t *sc, *xy; uint8_t* out; size_t first, count; } job_t;
static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  for (size_t i = j->first; i < j->first + j->count; i++) lane(j->sc + 32 * i, j->xy ? j->xy + 64 * i : NULL, j->out + 33 * i);
  return NULL;
}
/* n lanes of: Scalar::from_bytes, from_affine | generator, multiply, to_affine, to_bytes */
void oref_p256_batch_mul(const uint8_t* scalars_be, const uint8_t* points_xy_be, size_t n, uint8_t* out33, int threads) {
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
/* probes for the tests: op 0 mul, 1 square, 2 add, 3 sub, 4 neg, 5 invert on raw limbs (little-endian u64 x 4) */
void oref_p256_field_op(int op, const u64* a, const u64* b, u64* out) {
  fe x, y, r;
  memcpy(x.l, a, 32); memcpy(y.l, b, 32);
  r = op == 0 ? fe_mul(x, y) : op == 1 ? fe_square(x) : op == 2 ? fe_add(x, y) : op == 3 ? fe_sub(x, y) : op == 4 ? fe_neg(x) : fe_invert(x);
  memcpy(out, r.l, 32);
}
