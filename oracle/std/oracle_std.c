/* oracle_std.c -- TEST INFRASTRUCTURE: CPU oracle of the batched scalar-multiplication path.
 *
 * Only tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py may
 * call this.  The product (forge-ec_b200/) never links, imports or executes it.
 *
 * What it states: the behaviour forge-ec SPECIFIES for the path ("oracle_std"),
 * written the way the reference is written -- one point at a time, 4 x u64 limbs with u128
 * products, MSB-first double-and-add, one Fermat inversion per result -- but with correct field
 * arithmetic (generic 4-limb Montgomery multiplication for all three primes):
 *   w_mul            Curve::multiply + to_affine      forge-ec-curves/src/secp256k1.rs:2635-2692, 1342-1363
 *                                                      forge-ec-curves/src/p256.rs:2120-2156, 1835-1857
 *   jac_dbl/jac_add  ProjectivePoint::double / Add     secp256k1.rs:1502-1540, 1447-1497; p256.rs:1869-1912, 1941-2006
 *   f_inv            FieldElement::invert (a^(p-2))    secp256k1.rs:599-632; p256.rs:343-370; ed25519.rs:410-431
 *   scalar parse     Scalar::from_bytes (BE, < n)      secp256k1.rs:2271-2297; p256.rs:1041-1055
 *   encoders         PointAffine::to_bytes / CompressedPoint::from_affine / UncompressedPoint::from_affine
 *                                                      secp256k1.rs:875-894; forge-ec-encoding/src/point.rs:39-67, 186-211
 *   decode           CompressedPoint::to_affine        forge-ec-encoding/src/point.rs:70-144
 *   ostd_x25519      x25519(scalar,u)                  forge-ec-curves/src/curve25519.rs:1624-1716 (RFC 7748 section 5)
 *   ostd_ed25519     Ed25519Signature::derive_public_key  forge-ec-signature/src/eddsa.rs:450-509 (RFC 8032 5.1.5)
 *                    with Ed25519::multiply / Add      forge-ec-curves/src/ed25519.rs:2062-2097, 1864-1927
 *   sha512           the `sha2` crate the reference calls (sha2 0.10.9, eddsa.rs:464); FIPS 180-4
 *
 * Pinning: tests/test_oracle.py checks this file against every known-answer value the reference's
 * tests quote for the path, against oracle/pymodel.py and against
 * OpenSSL (`cryptography`) and libsodium (`PyNaCl`) on random inputs; fixtures in tests/golden/.
 * The reference itself cannot be built here (Rust, no toolchain) and its literal limb code does not
 * compute these functions; oracle/ref restates that literal code separately.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint64_t u64;
typedef unsigned __int128 u128;

/* ------------------------------------------------------------------ 256-bit helpers */
static int cmp4(const u64* a, const u64* b) {
  for (int i = 3; i >= 0; i--) {
    if (a[i] < b[i]) return -1;
    if (a[i] > b[i]) return 1;
  }
  return 0;
}
static u64 add4(u64* r, const u64* a, const u64* b) {
  u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)a[i] + b[i]; r[i] = (u64)c; c >>= 64; }
  return (u64)c;
}
static u64 sub4(u64* r, const u64* a, const u64* b) {
  u64 bw = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a[i] - b[i] - bw;
    r[i] = (u64)d;
    bw = (u64)(d >> 64) & 1;
  }
  return bw;
}
static int is_zero4(const u64* a) { return (a[0] | a[1] | a[2] | a[3]) == 0; }
static void set4(u64* r, u64 v) { r[0] = v; r[1] = r[2] = r[3] = 0; }
static void cpy4(u64* r, const u64* a) { memcpy(r, a, 32); }
static void from_be(u64* r, const uint8_t* b) {
  for (int i = 0; i < 4; i++) { u64 v = 0; for (int j = 0; j < 8; j++) v = (v << 8) | b[8 * (3 - i) + j]; r[i] = v; }
}
static void to_be(uint8_t* b, const u64* a) {
  for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) b[8 * (3 - i) + j] = (uint8_t)(a[i] >> (56 - 8 * j));
}
static void from_le(u64* r, const uint8_t* b) {
  for (int i = 0; i < 4; i++) { u64 v = 0; for (int j = 7; j >= 0; j--) v = (v << 8) | b[8 * i + j]; r[i] = v; }
}
static void to_le(uint8_t* b, const u64* a) {
  for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) b[8 * i + j] = (uint8_t)(a[i] >> (8 * j));
}
static int bit4(const u64* a, int i) { return (int)((a[i >> 6] >> (i & 63)) & 1); }

/* ------------------------------------------------------------------ generic Montgomery field */
typedef struct { u64 p[4], r2[4], one[4], n0; } field_t;

static void f_add(const field_t* f, u64* r, const u64* a, const u64* b) {
  u64 t[4];
  u64 c = add4(t, a, b);
  if (c || cmp4(t, f->p) >= 0) sub4(t, t, f->p);
  cpy4(r, t);
}
static void f_sub(const field_t* f, u64* r, const u64* a, const u64* b) {
  u64 t[4];
  if (sub4(t, a, b)) add4(t, t, f->p);
  cpy4(r, t);
}
static void f_mul(const field_t* f, u64* r, const u64* a, const u64* b) {
  u64 t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 uv;
    u64 c = 0;
    for (int j = 0; j < 4; j++) { uv = (u128)a[j] * b[i] + t[j] + c; t[j] = (u64)uv; c = (u64)(uv >> 64); }
    uv = (u128)t[4] + c; t[4] = (u64)uv; t[5] = (u64)(uv >> 64);
    u64 m = t[0] * f->n0;
    uv = (u128)m * f->p[0] + t[0]; c = (u64)(uv >> 64);
    for (int j = 1; j < 4; j++) { uv = (u128)m * f->p[j] + t[j] + c; t[j - 1] = (u64)uv; c = (u64)(uv >> 64); }
    uv = (u128)t[4] + c; t[3] = (u64)uv; t[4] = t[5] + (u64)(uv >> 64);
  }
  if (t[4] || cmp4(t, f->p) >= 0) sub4(t, t, f->p);
  cpy4(r, t);
}
static void f_sqr(const field_t* f, u64* r, const u64* a) { f_mul(f, r, a, a); }
static void f_to(const field_t* f, u64* r, const u64* a) { f_mul(f, r, a, f->r2); }
static void f_from(const field_t* f, u64* r, const u64* a) { u64 o[4]; set4(o, 1); f_mul(f, r, a, o); }
static void f_pow(const field_t* f, u64* r, const u64* a, const u64* e) {
  u64 acc[4], base[4];
  cpy4(acc, f->one); cpy4(base, a);
  for (int i = 255; i >= 0; i--) { f_sqr(f, acc, acc); if (bit4(e, i)) f_mul(f, acc, acc, base); }
  cpy4(r, acc);
}
static void f_inv(const field_t* f, u64* r, const u64* a) {
  u64 e[4], two[4]; set4(two, 2); sub4(e, f->p, two);
  f_pow(f, r, a, e);
}
static void f_neg(const field_t* f, u64* r, const u64* a) { u64 z[4]; set4(z, 0); f_sub(f, r, z, a); }
static void f_init(field_t* f, const u64 p[4]) {
  cpy4(f->p, p);
  u64 inv = 1;
  for (int i = 0; i < 7; i++) inv *= 2 - p[0] * inv;
  f->n0 = (u64)0 - inv;
  u64 x[4]; set4(x, 1);
  for (int i = 0; i < 512; i++) {
    u64 c = add4(x, x, x);
    if (c || cmp4(x, p) >= 0) sub4(x, x, p);
    if (i == 255) cpy4(f->one, x);
  }
  cpy4(f->r2, x);
}
static void f_set_ui(const field_t* f, u64* r, u64 v) { u64 t[4]; set4(t, v); f_to(f, r, t); }

/* ------------------------------------------------------------------ short Weierstrass */
typedef struct { field_t F; u64 n[4]; u64 a[4], b[4], gx[4], gy[4]; int a_is_zero; } wcurve_t;
typedef struct { u64 X[4], Y[4], Z[4]; } jac_t;
static wcurve_t W[2];
static field_t F25519;
static u64 ED_D[4], ED_BX[4], ED_BY[4], X_A24[4];
static int g_init_done = 0;

static void hex4(u64* r, const char* h) {  /* 64 hex digits, big-endian */
  uint8_t b[32];
  for (int i = 0; i < 32; i++) {
    unsigned v = 0;
    for (int j = 0; j < 2; j++) { char c = h[2 * i + j]; v = v * 16 + (unsigned)(c <= '9' ? c - '0' : (c | 32) - 'a' + 10); }
    b[i] = (uint8_t)v;
  }
  from_be(r, b);
}
static void ostd_init_once(void) {
  if (g_init_done) return;
  u64 t[4];
  /* SEC2 secp256k1 (values the reference quotes at secp256k1.rs:22-28, 2610-2622) */
  hex4(t, "FFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEFFFFFC2F"); f_init(&W[0].F, t);
  hex4(W[0].n, "FFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141");
  set4(W[0].a, 0); f_set_ui(&W[0].F, W[0].b, 7); W[0].a_is_zero = 1;
  hex4(t, "79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798"); f_to(&W[0].F, W[0].gx, t);
  hex4(t, "483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8"); f_to(&W[0].F, W[0].gy, t);
  /* FIPS 186-4 P-256 (p256.rs:18-33, 2094-2106) */
  hex4(t, "FFFFFFFF00000001000000000000000000000000FFFFFFFFFFFFFFFFFFFFFFFF"); f_init(&W[1].F, t);
  hex4(W[1].n, "FFFFFFFF00000000FFFFFFFFFFFFFFFFBCE6FAADA7179E84F3B9CAC2FC632551");
  { u64 three[4]; f_set_ui(&W[1].F, three, 3); f_neg(&W[1].F, W[1].a, three); } W[1].a_is_zero = 0;
  hex4(t, "5AC635D8AA3A93E7B3EBBD55769886BC651D06B0CC53B0F63BCE3C3E27D2604B"); f_to(&W[1].F, W[1].b, t);
  hex4(t, "6B17D1F2E12C4247F8BCE6E563A440F277037D812DEB33A0F4A13945D898C296"); f_to(&W[1].F, W[1].gx, t);
  hex4(t, "4FE342E2FE1A7F9B8EE7EB4A7C0F9E162BCE33576B315ECECBB6406837BF51F5"); f_to(&W[1].F, W[1].gy, t);
  /* 2^255 - 19 */
  hex4(t, "7FFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFED"); f_init(&F25519, t);
  { u64 a[4], b[4], bi[4];
    f_set_ui(&F25519, a, 121665); f_set_ui(&F25519, b, 121666);
    f_inv(&F25519, bi, b); f_mul(&F25519, a, a, bi); f_neg(&F25519, ED_D, a);        /* d = -121665/121666 */
    f_set_ui(&F25519, X_A24, 121665); }
  hex4(t, "216936D3CD6E53FEC0A4E231FDD6DC5C692CC7609525A7B2C9562D608F25D51A"); f_to(&F25519, ED_BX, t);
  hex4(t, "6666666666666666666666666666666666666666666666666666666666666658"); f_to(&F25519, ED_BY, t);
  g_init_done = 1;
}

static void jac_dbl(const wcurve_t* c, jac_t* r, const jac_t* p) {
  const field_t* f = &c->F;
  if (is_zero4(p->Z) || is_zero4(p->Y)) { cpy4(r->X, f->one); cpy4(r->Y, f->one); set4(r->Z, 0); return; }
  u64 yy[4], s[4], m[4], t[4], zz[4], x3[4], y3[4], z3[4];
  f_sqr(f, yy, p->Y);
  f_mul(f, s, p->X, yy); f_add(f, s, s, s); f_add(f, s, s, s);           /* S = 4 X Y^2 */
  f_sqr(f, m, p->X); f_add(f, t, m, m); f_add(f, m, t, m);               /* 3 X^2 */
  if (!c->a_is_zero) { f_sqr(f, zz, p->Z); f_sqr(f, zz, zz); f_mul(f, zz, zz, c->a); f_add(f, m, m, zz); }
  f_sqr(f, x3, m); f_sub(f, x3, x3, s); f_sub(f, x3, x3, s);
  f_sub(f, t, s, x3); f_mul(f, y3, m, t);
  f_sqr(f, t, yy); f_add(f, t, t, t); f_add(f, t, t, t); f_add(f, t, t, t);  /* 8 Y^4 */
  f_sub(f, y3, y3, t);
  f_mul(f, z3, p->Y, p->Z); f_add(f, z3, z3, z3);
  cpy4(r->X, x3); cpy4(r->Y, y3); cpy4(r->Z, z3);
}
static void jac_add(const wcurve_t* c, jac_t* r, const jac_t* p, const jac_t* q) {
  const field_t* f = &c->F;
  if (is_zero4(p->Z)) { *r = *q; return; }
  if (is_zero4(q->Z)) { *r = *p; return; }
  u64 z1z1[4], z2z2[4], u1[4], u2[4], s1[4], s2[4], h[4], rr[4], hh[4], hhh[4], v[4], t[4], x3[4], y3[4], z3[4];
  f_sqr(f, z1z1, p->Z); f_sqr(f, z2z2, q->Z);
  f_mul(f, u1, p->X, z2z2); f_mul(f, u2, q->X, z1z1);
  f_mul(f, s1, p->Y, q->Z); f_mul(f, s1, s1, z2z2);
  f_mul(f, s2, q->Y, p->Z); f_mul(f, s2, s2, z1z1);
  if (cmp4(u1, u2) == 0) {
    if (cmp4(s1, s2) == 0) { jac_dbl(c, r, p); return; }
    cpy4(r->X, f->one); cpy4(r->Y, f->one); set4(r->Z, 0); return;
  }
  f_sub(f, h, u2, u1); f_sub(f, rr, s2, s1);
  f_sqr(f, hh, h); f_mul(f, hhh, hh, h); f_mul(f, v, u1, hh);
  f_sqr(f, x3, rr); f_sub(f, x3, x3, hhh); f_sub(f, x3, x3, v); f_sub(f, x3, x3, v);
  f_sub(f, t, v, x3); f_mul(f, y3, rr, t); f_mul(f, t, s1, hhh); f_sub(f, y3, y3, t);
  f_mul(f, z3, p->Z, q->Z); f_mul(f, z3, z3, h);
  cpy4(r->X, x3); cpy4(r->Y, y3); cpy4(r->Z, z3);
}
static int w_on_curve(const wcurve_t* c, const u64* xm, const u64* ym) {
  const field_t* f = &c->F;
  u64 l[4], r[4], t[4];
  f_sqr(f, l, ym);
  f_sqr(f, r, xm); f_mul(f, r, r, xm);
  f_mul(f, t, c->a, xm); f_add(f, r, r, t); f_add(f, r, r, c->b);
  return cmp4(l, r) == 0;
}
/* k (plain integer) times affine (x,y) in Montgomery form; returns 0 if the result is infinity */
static int w_mul(const wcurve_t* c, const u64* k, const u64* xm, const u64* ym, u64* ox, u64* oy) {
  const field_t* f = &c->F;
  jac_t q, b;
  cpy4(q.X, f->one); cpy4(q.Y, f->one); set4(q.Z, 0);
  cpy4(b.X, xm); cpy4(b.Y, ym); cpy4(b.Z, f->one);
  for (int i = 255; i >= 0; i--) {
    jac_dbl(c, &q, &q);
    if (bit4(k, i)) jac_add(c, &q, &q, &b);
  }
  if (is_zero4(q.Z)) return 0;
  u64 zi[4], zi2[4];
  f_inv(f, zi, q.Z); f_sqr(f, zi2, zi);
  f_mul(f, ox, q.X, zi2); f_mul(f, zi2, zi2, zi); f_mul(f, oy, q.Y, zi2);
  return 1;
}
static int out_size(int fmt) { return fmt == 0 ? 33 : ((fmt == 1 || fmt == 4) ? 65 : (fmt == 3 ? 32 : 64)); }

/* one lane of fec_batch_mul; returns the status byte */
static int w_lane(int curve, const uint8_t* sc, const uint8_t* pt, int pfmt, int ofmt, uint8_t* out) {
  const wcurve_t* c = &W[curve];
  const field_t* f = &c->F;
  const int osz = out_size(ofmt);
  memset(out, 0, (size_t)osz);
  u64 k[4], x[4], y[4], xm[4], ym[4];
  from_be(k, sc);
  if (cmp4(k, c->n) >= 0) return 1;
  if (pt == NULL) {
    cpy4(xm, c->gx); cpy4(ym, c->gy);
  } else if (pfmt == 0) {
    from_be(x, pt); from_be(y, pt + 32);
    if (is_zero4(x) && is_zero4(y)) return 3;                 /* the identity as an input point (all-zero record, point.rs:41-47): k * O = O */
    if (cmp4(x, f->p) >= 0 || cmp4(y, f->p) >= 0) return 2;
    f_to(f, xm, x); f_to(f, ym, y);
    if (!w_on_curve(c, xm, ym)) return 2;
  } else {
    if (pt[0] == 0) return 3;                                 /* identity marker, point.rs:82-87 */
    if (pt[0] != 2 && pt[0] != 3) return 2;
    from_be(x, pt + 1);
    if (cmp4(x, f->p) >= 0) return 2;
    f_to(f, xm, x);
    u64 rhs[4], t[4], e[4], one[4];
    f_sqr(f, rhs, xm); f_mul(f, rhs, rhs, xm); f_mul(f, t, c->a, xm); f_add(f, rhs, rhs, t); f_add(f, rhs, rhs, c->b);
    set4(one, 1); add4(e, f->p, one);                         /* (p+1)/4: p+1 does not overflow 2^256 for these primes */
    for (int i = 0; i < 3; i++) e[i] = (e[i] >> 2) | (e[i + 1] << 62);
    e[3] >>= 2;
    f_pow(f, ym, rhs, e);
    f_sqr(f, t, ym);
    if (cmp4(t, rhs) != 0) return 2;
    f_from(f, y, ym);
    if ((int)(y[0] & 1) != (pt[0] & 1)) f_neg(f, ym, ym);
  }
  u64 rx[4], ry[4];
  if (!w_mul(c, k, xm, ym, rx, ry)) return 3;
  f_from(f, x, rx); f_from(f, y, ry);
  if (ofmt == 0) { out[0] = (uint8_t)(2 | (y[0] & 1)); to_be(out + 1, x); }
  else if (ofmt == 1) { out[0] = 4; to_be(out + 1, x); to_be(out + 33, y); }
  else if (ofmt == 4) { out[0] = (uint8_t)(6 | (y[0] & 1)); to_be(out + 1, x); to_be(out + 33, y); }  /* PointFormat::Hybrid, secp256k1.rs:1048-1066 */
  else if (ofmt == 3) { to_be(out, x); }                                                              /* x only: the ECDH shared secret */
  else { to_be(out, x); to_be(out + 32, y); }
  return 0;
}

/* ------------------------------------------------------------------ X25519 (RFC 7748 section 5) */
static void x25519_lane(const uint8_t* ks, const uint8_t* us, uint8_t* out) {
  const field_t* f = &F25519;
  uint8_t kb[32], ub[32];
  memcpy(kb, ks, 32); kb[0] &= 248; kb[31] &= 127; kb[31] |= 64;
  memcpy(ub, us, 32); ub[31] &= 127;
  u64 k[4], u[4], x1[4], x2[4], z2[4], x3[4], z3[4];
  from_le(k, kb); from_le(u, ub);
  if (cmp4(u, f->p) >= 0) sub4(u, u, f->p);
  f_to(f, x1, u);
  cpy4(x2, f->one); set4(z2, 0); cpy4(x3, x1); cpy4(z3, f->one);
  int swap = 0;
  for (int t = 254; t >= 0; t--) {
    int kt = bit4(k, t);
    swap ^= kt;
    if (swap) { u64 s[4]; cpy4(s, x2); cpy4(x2, x3); cpy4(x3, s); cpy4(s, z2); cpy4(z2, z3); cpy4(z3, s); }
    swap = kt;
    u64 A[4], AA[4], B[4], BB[4], E[4], C[4], D[4], DA[4], CB[4], t0[4];
    f_add(f, A, x2, z2); f_sqr(f, AA, A); f_sub(f, B, x2, z2); f_sqr(f, BB, B); f_sub(f, E, AA, BB);
    f_add(f, C, x3, z3); f_sub(f, D, x3, z3); f_mul(f, DA, D, A); f_mul(f, CB, C, B);
    f_add(f, t0, DA, CB); f_sqr(f, x3, t0);
    f_sub(f, t0, DA, CB); f_sqr(f, t0, t0); f_mul(f, z3, x1, t0);
    f_mul(f, x2, AA, BB);
    f_mul(f, t0, X_A24, E); f_add(f, t0, AA, t0); f_mul(f, z2, E, t0);
  }
  if (swap) { u64 s[4]; cpy4(s, x2); cpy4(x2, x3); cpy4(x3, s); cpy4(s, z2); cpy4(z2, z3); cpy4(z3, s); }
  u64 zi[4], r[4];
  f_inv(f, zi, z2); f_mul(f, r, x2, zi); f_from(f, r, r);
  to_le(out, r);
}

/* ------------------------------------------------------------------ SHA-512 (FIPS 180-4) */
static const u64 SHA512_K[80] = {
  0x428A2F98D728AE22ULL, 0x7137449123EF65CDULL, 0xB5C0FBCFEC4D3B2FULL, 0xE9B5DBA58189DBBCULL,
  0x3956C25BF348B538ULL, 0x59F111F1B605D019ULL, 0x923F82A4AF194F9BULL, 0xAB1C5ED5DA6D8118ULL,
  0xD807AA98A3030242ULL, 0x12835B0145706FBEULL, 0x243185BE4EE4B28CULL, 0x550C7DC3D5FFB4E2ULL,
  0x72BE5D74F27B896FULL, 0x80DEB1FE3B1696B1ULL, 0x9BDC06A725C71235ULL, 0xC19BF174CF692694ULL,
  0xE49B69C19EF14AD2ULL, 0xEFBE4786384F25E3ULL, 0x0FC19DC68B8CD5B5ULL, 0x240CA1CC77AC9C65ULL,
  0x2DE92C6F592B0275ULL, 0x4A7484AA6EA6E483ULL, 0x5CB0A9DCBD41FBD4ULL, 0x76F988DA831153B5ULL,
  0x983E5152EE66DFABULL, 0xA831C66D2DB43210ULL, 0xB00327C898FB213FULL, 0xBF597FC7BEEF0EE4ULL,
  0xC6E00BF33DA88FC2ULL, 0xD5A79147930AA725ULL, 0x06CA6351E003826FULL, 0x142929670A0E6E70ULL,
  0x27B70A8546D22FFCULL, 0x2E1B21385C26C926ULL, 0x4D2C6DFC5AC42AEDULL, 0x53380D139D95B3DFULL,
  0x650A73548BAF63DEULL, 0x766A0ABB3C77B2A8ULL, 0x81C2C92E47EDAEE6ULL, 0x92722C851482353BULL,
  0xA2BFE8A14CF10364ULL, 0xA81A664BBC423001ULL, 0xC24B8B70D0F89791ULL, 0xC76C51A30654BE30ULL,
  0xD192E819D6EF5218ULL, 0xD69906245565A910ULL, 0xF40E35855771202AULL, 0x106AA07032BBD1B8ULL,
  0x19A4C116B8D2D0C8ULL, 0x1E376C085141AB53ULL, 0x2748774CDF8EEB99ULL, 0x34B0BCB5E19B48A8ULL,
  0x391C0CB3C5C95A63ULL, 0x4ED8AA4AE3418ACBULL, 0x5B9CCA4F7763E373ULL, 0x682E6FF3D6B2B8A3ULL,
  0x748F82EE5DEFB2FCULL, 0x78A5636F43172F60ULL, 0x84C87814A1F0AB72ULL, 0x8CC702081A6439ECULL,
  0x90BEFFFA23631E28ULL, 0xA4506CEBDE82BDE9ULL, 0xBEF9A3F7B2C67915ULL, 0xC67178F2E372532BULL,
  0xCA273ECEEA26619CULL, 0xD186B8C721C0C207ULL, 0xEADA7DD6CDE0EB1EULL, 0xF57D4F7FEE6ED178ULL,
  0x06F067AA72176FBAULL, 0x0A637DC5A2C898A6ULL, 0x113F9804BEF90DAEULL, 0x1B710B35131C471BULL,
  0x28DB77F523047D84ULL, 0x32CAAB7B40C72493ULL, 0x3C9EBE0A15C9BEBCULL, 0x431D67C49C100D4CULL,
  0x4CC5D4BECB3E42B6ULL, 0x597F299CFC657E2AULL, 0x5FCB6FAB3AD6FAECULL, 0x6C44198C4A475817ULL};
static u64 rotr(u64 x, int n) { return (x >> n) | (x << (64 - n)); }
static void sha512(const uint8_t* msg, size_t len, uint8_t digest[64]) {
  u64 h[8] = {0x6A09E667F3BCC908ULL, 0xBB67AE8584CAA73BULL, 0x3C6EF372FE94F82BULL, 0xA54FF53A5F1D36F1ULL, 0x510E527FADE682D1ULL, 0x9B05688C2B3E6C1FULL, 0x1F83D9ABFB41BD6BULL, 0x5BE0CD19137E2179ULL};
  size_t total = ((len + 17 + 127) / 128) * 128;
  uint8_t* buf = (uint8_t*)calloc(total, 1);
  memcpy(buf, msg, len);
  buf[len] = 0x80;
  u64 bits = (u64)len * 8;
  for (int i = 0; i < 8; i++) buf[total - 1 - i] = (uint8_t)(bits >> (8 * i));
  for (size_t off = 0; off < total; off += 128) {
    u64 w[80];
    for (int i = 0; i < 16; i++) { u64 v = 0; for (int j = 0; j < 8; j++) v = (v << 8) | buf[off + 8 * i + j]; w[i] = v; }
    for (int i = 16; i < 80; i++) {
      u64 s0 = rotr(w[i - 15], 1) ^ rotr(w[i - 15], 8) ^ (w[i - 15] >> 7);
      u64 s1 = rotr(w[i - 2], 19) ^ rotr(w[i - 2], 61) ^ (w[i - 2] >> 6);
      w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    u64 a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 80; i++) {
      u64 t1 = hh + (rotr(e, 14) ^ rotr(e, 18) ^ rotr(e, 41)) + ((e & f) ^ (~e & g)) + SHA512_K[i] + w[i];
      u64 t2 = (rotr(a, 28) ^ rotr(a, 34) ^ rotr(a, 39)) + ((a & b) ^ (a & c) ^ (b & c));
      hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
  }
  free(buf);
  for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) digest[8 * i + j] = (uint8_t)(h[i] >> (56 - 8 * j));
}

/* ------------------------------------------------------------------ Ed25519 public key */
typedef struct { u64 X[4], Y[4], Z[4], T[4]; } ext_t;
static void ed_add(ext_t* r, const ext_t* p, const ext_t* q) {   /* add-2008-hwcd-3, a = -1 */
  const field_t* f = &F25519;
  u64 A[4], B[4], C[4], D[4], E[4], Fv[4], G[4], H[4], t[4];
  f_sub(f, A, p->Y, p->X); f_sub(f, t, q->Y, q->X); f_mul(f, A, A, t);
  f_add(f, B, p->Y, p->X); f_add(f, t, q->Y, q->X); f_mul(f, B, B, t);
  f_mul(f, C, p->T, q->T); f_mul(f, C, C, ED_D); f_add(f, C, C, C);
  f_mul(f, D, p->Z, q->Z); f_add(f, D, D, D);
  f_sub(f, E, B, A); f_sub(f, Fv, D, C); f_add(f, G, D, C); f_add(f, H, B, A);
  f_mul(f, r->X, E, Fv); f_mul(f, r->Y, G, H); f_mul(f, r->T, E, H); f_mul(f, r->Z, Fv, G);
}
static void ed25519_lane(const uint8_t* in, int is_seed, uint8_t* out) {
  const field_t* f = &F25519;
  uint8_t sb[32];
  if (is_seed) {
    uint8_t dg[64];
    sha512(in, 32, dg);
    memcpy(sb, dg, 32); sb[0] &= 248; sb[31] &= 127; sb[31] |= 64;
  } else {
    memcpy(sb, in, 32);
  }
  u64 s[4];
  from_le(s, sb);
  ext_t q, b;
  set4(q.X, 0); cpy4(q.Y, f->one); cpy4(q.Z, f->one); set4(q.T, 0);
  cpy4(b.X, ED_BX); cpy4(b.Y, ED_BY); cpy4(b.Z, f->one); f_mul(f, b.T, ED_BX, ED_BY);
  for (int i = 0; i < 256; i++) {                /* LSB-first double-and-add, as ed25519.rs:2062-2097 */
    if (bit4(s, i)) ed_add(&q, &q, &b);
    ed_add(&b, &b, &b);
  }
  u64 zi[4], x[4], y[4];
  f_inv(f, zi, q.Z); f_mul(f, x, q.X, zi); f_mul(f, y, q.Y, zi);
  f_from(f, x, x); f_from(f, y, y);
  y[3] |= (x[0] & 1) << 63;
  to_le(out, y);
}

/* ------------------------------------------------------------------ field microbench model */
static void field_lane(int prime, int op, int iters, const uint8_t* a, const uint8_t* b, uint8_t* out) {
  const field_t* f = prime == 0 ? &W[0].F : (prime == 1 ? &W[1].F : &F25519);
  u64 x[4], y[4];
  from_le(x, a); from_le(y, b);
  f_to(f, x, x); f_to(f, y, y);
  for (int i = 0; i < iters; i++) {
    if (op == 0) f_mul(f, x, x, y);
    else if (op == 1) f_sqr(f, x, x);
    else { f_mul(f, x, x, y); f_sqr(f, x, x); }
  }
  f_from(f, x, x);
  to_le(out, x);
}

/* ------------------------------------------------------------------ batch drivers (contiguous shards, pthreads) */
typedef struct {
  int kind, curve, pfmt, ofmt, flag, iters;
  const uint8_t *a, *b; uint8_t *out, *status;
  size_t first, count;
} job_t;
static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  for (size_t i = j->first; i < j->first + j->count; i++) {
    if (j->kind == 0) {
      const int psz = j->pfmt == 0 ? 64 : 33, osz = out_size(j->ofmt);
      int st = w_lane(j->curve, j->a + 32 * i, j->b ? j->b + (size_t)psz * i : NULL, j->pfmt, j->ofmt, j->out + (size_t)osz * i);
      if (j->status) j->status[i] = (uint8_t)st;
    } else if (j->kind == 1) {
      static const uint8_t nine[32] = {9};
      x25519_lane(j->a + 32 * i, j->b ? j->b + 32 * i : nine, j->out + 32 * i);
    } else if (j->kind == 2) {
      ed25519_lane(j->a + 32 * i, j->flag, j->out + 32 * i);
    } else {
      field_lane(j->curve, j->flag, j->iters, j->a + 32 * i, j->b + 32 * i, j->out + 32 * i);
    }
  }
  return NULL;
}
static void run(job_t base, size_t n, int threads) {
  ostd_init_once();
  if (threads < 1) threads = 1;
  if ((size_t)threads > n) threads = n ? (int)n : 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
  job_t* jobs = (job_t*)malloc(sizeof(job_t) * (size_t)threads);
  for (int t = 0; t < threads; t++) {
    jobs[t] = base;
    jobs[t].first = n * (size_t)t / (size_t)threads;
    jobs[t].count = n * (size_t)(t + 1) / (size_t)threads - jobs[t].first;
    if (threads == 1) worker(&jobs[t]); else pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  if (threads > 1) for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
  free(th); free(jobs);
}

void ostd_batch_mul(int curve, const uint8_t* scalars, const uint8_t* points, int points_fmt, size_t n, int out_fmt,
                    uint8_t* out, uint8_t* status, int threads) {
  job_t j = {0, curve, points_fmt, out_fmt, 0, 0, scalars, points, out, status, 0, 0};
  run(j, n, threads);
}
/* Curve::multi_scalar_multiply (forge-ec-core/src/lib.rs:934-951): the products multiply(points[i], scalars[i]) added one by one.
 * out: ONE record of out_fmt, status: ONE byte -- 1 / 2 if some lane has an out-of-range scalar / an invalid point (all-zero
 * record), 3 if the sum is the identity.  ostd_msm_products adds products that ostd_batch_mul has already computed (n x 64 B
 * X || Y + n status bytes); ostd_msm computes them with `threads` workers first.  The additions run in order. */
void ostd_msm_products(int curve, const uint8_t* xy, const uint8_t* st, size_t n, int out_fmt, uint8_t* out, uint8_t* status) {
  ostd_init_once();
  const wcurve_t* c = &W[curve];
  const field_t* f = &c->F;
  memset(out, 0, (size_t)out_size(out_fmt));
  int worst = 0;
  for (size_t i = 0; i < n; i++) if (st[i] == 1 || st[i] == 2) { if (worst != 1) worst = st[i]; }
  if (worst) { *status = (uint8_t)worst; return; }
  jac_t acc, t;
  cpy4(acc.X, f->one); cpy4(acc.Y, f->one); set4(acc.Z, 0);
  for (size_t i = 0; i < n; i++) {
    if (st[i] != 0) continue;                                   /* identity products add nothing */
    u64 x[4], y[4];
    from_be(x, xy + 64 * i); from_be(y, xy + 64 * i + 32);
    f_to(f, t.X, x); f_to(f, t.Y, y); cpy4(t.Z, f->one);
    jac_add(c, &acc, &acc, &t);
  }
  if (is_zero4(acc.Z)) { *status = 3; return; }
  u64 zi[4], zi2[4], ax[4], ay[4], x[4], y[4];
  f_inv(f, zi, acc.Z); f_sqr(f, zi2, zi);
  f_mul(f, ax, acc.X, zi2); f_mul(f, zi2, zi2, zi); f_mul(f, ay, acc.Y, zi2);
  f_from(f, x, ax); f_from(f, y, ay);
  if (out_fmt == 0) { out[0] = (uint8_t)(2 | (y[0] & 1)); to_be(out + 1, x); }
  else if (out_fmt == 1) { out[0] = 4; to_be(out + 1, x); to_be(out + 33, y); }
  else if (out_fmt == 4) { out[0] = (uint8_t)(6 | (y[0] & 1)); to_be(out + 1, x); to_be(out + 33, y); }
  else if (out_fmt == 3) { to_be(out, x); }
  else { to_be(out, x); to_be(out + 32, y); }
  *status = 0;
}
void ostd_msm(int curve, const uint8_t* scalars, const uint8_t* points, int points_fmt, size_t n, int out_fmt, uint8_t* out,
              uint8_t* status, int threads) {
  uint8_t* xy = (uint8_t*)malloc(n * 64 + 1);
  uint8_t* st = (uint8_t*)malloc(n + 1);
  ostd_batch_mul(curve, scalars, points, points_fmt, n, 2, xy, st, threads);
  ostd_msm_products(curve, xy, st, n, out_fmt, out, status);
  free(xy); free(st);
}
/* Standalone point decoding / validation (CompressedPoint::to_affine, forge-ec-encoding/src/point.rs:70-144; Curve::validate_point,
 * forge-ec-core/src/lib.rs:905-923): k = 1 times the decoded point is the point itself, so this is w_lane with the scalar 1 --
 * status 2 for a rejected record, 3 for the identity record. */
void ostd_batch_point_decode(int curve, const uint8_t* points, int points_fmt, size_t n, int out_fmt, uint8_t* out, uint8_t* status,
                             int threads) {
  uint8_t* one = (uint8_t*)calloc(n ? n : 1, 32);
  for (size_t i = 0; i < n; i++) one[32 * i + 31] = 1;
  ostd_batch_mul(curve, one, points, points_fmt, n, out_fmt, out, status, threads);
  free(one);
}
void ostd_batch_ed25519_pubkey(const uint8_t* in, int input_is_seed, size_t n, uint8_t* out, int threads);
/* Signing-side nonce commitment.  mode 0 ECDSA (forge-ec-signature/src/ecdsa.rs:108-131): R = k G (33 B SEC1
 * compressed), r = x(R) as a scalar, status 1 where Scalar::from_bytes(x) fails (x >= n) and 3 for r = 0; mode 1 BIP-340
 * (schnorr.rs:372-385): x-only R (32 B), k' = n - k when y(R) is odd; mode 2 EdDSA (RFC 8032 5.1.6): r = the 64-byte little-endian
 * value mod L, R = [r]B.  out_r: 32 B per lane (big-endian for the Weierstrass modes, little-endian for EdDSA). */
static const u64 ED_L[4] = {0x5812631A5CF5D3EDull, 0x14DEF9DEA2F79CD6ull, 0, 0x1000000000000000ull};
static void sc_reduce64(const uint8_t* in, uint8_t* out_le) {   /* byte-wise Horner: r = (256 r + b) mod L */
  u64 r[4] = {0, 0, 0, 0};
  for (int i = 63; i >= 0; i--) {
    u64 v[5];
    v[4] = r[3] >> 56;
    v[3] = (r[3] << 8) | (r[2] >> 56); v[2] = (r[2] << 8) | (r[1] >> 56); v[1] = (r[1] << 8) | (r[0] >> 56); v[0] = (r[0] << 8) | in[i];
    /* v < 2^261: subtract L while v >= L (at most 2^9 + 1 times would be slow): q = v >> 252, v -= q L, then fix the sign */
    u64 q = (v[3] >> 60) | (v[4] << 4);
    u64 t[4] = {v[0], v[1], v[2], v[3] & 0x0FFFFFFFFFFFFFFFull};
    u128 m0 = (u128)q * ED_L[0], m1 = (u128)q * ED_L[1] + (u64)(m0 >> 64);
    u64 s[4] = {(u64)m0, (u64)m1, (u64)(m1 >> 64), 0};
    u64 d[4];
    if (sub4(d, t, s)) add4(d, d, ED_L);
    memcpy(r, d, 32);
  }
  to_le(out_le, r);
}
void ostd_batch_nonce_commit(int curve, int mode, const uint8_t* nonces, size_t n, uint8_t* out_R, uint8_t* out_r, uint8_t* status,
                             int threads) {
  ostd_init_once();
  if (mode == 2) {
    for (size_t i = 0; i < n; i++) sc_reduce64(nonces + 64 * i, out_r + 32 * i);
    ostd_batch_ed25519_pubkey(out_r, 0, n, out_R, threads);
    if (status) memset(status, 0, n);
    return;
  }
  const wcurve_t* c = &W[curve];
  uint8_t* xy = (uint8_t*)malloc(n * 64 + 1);
  uint8_t* st = (uint8_t*)malloc(n + 1);
  ostd_batch_mul(curve, nonces, NULL, 0, n, 2, xy, st, threads);
  const size_t rsz = mode == 0 ? 33 : 32;
  memset(out_R, 0, n * rsz); memset(out_r, 0, n * 32);
  for (size_t i = 0; i < n; i++) {
    int s = st[i];
    u64 x[4], y[4], k[4];
    from_be(x, xy + 64 * i); from_be(y, xy + 64 * i + 32); from_be(k, nonces + 32 * i);
    if (s == 0 && mode == 0) s = cmp4(x, c->n) >= 0 ? 1 : (is_zero4(x) ? 3 : 0);
    if (s == 0) {
      if (mode == 0) {
        out_R[33 * i] = (uint8_t)(2 | (y[0] & 1)); to_be(out_R + 33 * i + 1, x); to_be(out_r + 32 * i, x);
      } else {
        u64 nk[4];
        sub4(nk, c->n, k);
        to_be(out_R + 32 * i, x); to_be(out_r + 32 * i, (y[0] & 1) ? nk : k);
      }
    }
    if (status) status[i] = (uint8_t)s;
  }
  free(xy); free(st);
}

void ostd_batch_x25519(const uint8_t* scalars, const uint8_t* us, size_t n, uint8_t* out, int threads) {
  job_t j = {1, 0, 0, 0, 0, 0, scalars, us, out, NULL, 0, 0};
  run(j, n, threads);
}
void ostd_batch_ed25519_pubkey(const uint8_t* in, int input_is_seed, size_t n, uint8_t* out, int threads) {
  job_t j = {2, 0, 0, 0, input_is_seed, 0, in, NULL, out, NULL, 0, 0};
  run(j, n, threads);
}
void ostd_field_bench(int prime, int op, size_t lanes, int iters, const uint8_t* a, const uint8_t* b, uint8_t* out, int threads) {
  job_t j = {3, prime, 0, 0, op, iters, a, b, out, NULL, 0, 0};
  run(j, lanes, threads);
}
void ostd_sha512(const uint8_t* msg, size_t len, uint8_t* digest) { sha512(msg, len, digest); }
