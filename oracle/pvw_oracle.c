/* CPU oracle (C restatement) for the pvw-rs hot path -- TEST INFRASTRUCTURE ONLY.
 *
 * Nothing here is part of the product.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, as the checker or as the timed CPU baseline.
 *
 * PARITY UNPINNED at the fhe-math boundary: the reference (gnosisguild/pvw-rs) is Rust, its ring arithmetic
 * lives in the un-vendored crates fhe-math/fhe-util/fhe-traits 0.1.0-beta.7
 * (gnosisguild/fhe.rs#364335035e3a801573539274b2d3052d5b69098a), no Rust toolchain exists in this image and
 * the reference holds no known-answer vectors.  This file restates the published algorithms; it is checked
 * bit-for-bit against oracle/pvw_oracle.py (exact Python integers) and against the reference's behavioural
 * tests restated in tests/.  psi (the 2l-th primitive root per modulus) is an explicit input.
 *
 * Plain C: u64 residues, unsigned __int128 products reduced with `%`, multi-precision integers as
 * little-endian u64 words with schoolbook multiply and Knuth division.  Parallel structure mirrors the
 * reference's rayon use (OpenMP over dealers x parties: src/crypto/encryption.rs:177-200,277-283;
 * src/crypto/decryption.rs:257-263,312-322).
 *
 * Host layout everywhere = the reference's: a polynomial is u64[L][ell] row-major (fhe-math Array2<u64>,
 * src/params/parameters.rs:455-458); matrices are arrays of polynomials.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;

#define PVWO_MAX_L 64
#define PVWO_MAX_ELL 256
#define BW 160 /* capacity of a big integer in 64-bit words */

typedef struct {
  uint32_t n, k, ell, L;
  uint32_t nw;                 /* words per big-integer constant below */
  const uint64_t *moduli;      /* [L] */
  const uint64_t *psi;         /* [L] primitive 2*ell-th roots */
  const uint64_t *Q;           /* [nw] product of the moduli */
  const uint64_t *delta;       /* [nw] floor(Q^(1/ell))            parameters.rs:156 */
  const uint64_t *delta_pow;   /* [nw] delta^(ell-1)               parameters.rs:159-163 */
  const uint64_t *qhat;        /* [L][nw] Q / q_j */
  const uint64_t *qhat_inv;    /* [L] (Q/q_j)^-1 mod q_j */
  const uint64_t *gadget_rns;  /* [L][ell] delta^t mod q_j (power basis), parameters.rs:288-308 */
} pvwo_params;

/* ------------------------------------------------------------------------------------------------ */
/* modular arithmetic (fhe-math zq::Modulus semantics: canonical results)                            */
/* ------------------------------------------------------------------------------------------------ */
static inline uint64_t mulmod(uint64_t a, uint64_t b, uint64_t q) { return (uint64_t)((u128)a * b % q); }
/* Barrett form of the same product for a, b < q < 2^62 (fhe-math zq::Modulus::mul = 128-bit product + Barrett
 * `reduce_u128`): mu = floor(2^128 / q) as (mu_hi, mu_lo); the quotient estimate is short by at most 3. */
typedef struct { uint64_t q, mu_hi, mu_lo; } bmod;
static inline bmod bmod_init(uint64_t q) {
  bmod m; m.q = q;
  u128 mu = (~(u128)0) / q; /* q odd prime: floor((2^128-1)/q) == floor(2^128/q) */
  m.mu_hi = (uint64_t)(mu >> 64); m.mu_lo = (uint64_t)mu;
  return m;
}
static inline uint64_t mulmod_b(uint64_t a, uint64_t b, const bmod *m) {
  u128 x = (u128)a * b;
  uint64_t h = (uint64_t)(x >> 64), lo = (uint64_t)x;
  uint64_t qh = h * m->mu_hi + (uint64_t)(((u128)h * m->mu_lo) >> 64) + (uint64_t)(((u128)lo * m->mu_hi) >> 64);
  uint64_t r = lo - qh * m->q;
  while (r >= m->q) r -= m->q;
  return r;
}
static inline uint64_t addmod(uint64_t a, uint64_t b, uint64_t q) { uint64_t s = a + b; return s >= q ? s - q : s; }
static inline uint64_t submod(uint64_t a, uint64_t b, uint64_t q) { return a >= b ? a - b : a + q - b; }
static uint64_t powmod(uint64_t a, uint64_t e, uint64_t q) {
  uint64_t r = 1 % q;
  a %= q;
  while (e) { if (e & 1) r = mulmod(r, a, q); a = mulmod(a, a, q); e >>= 1; }
  return r;
}
/* i64 -> canonical residue: ((x % q) + q) % q, parameters.rs:437-452 / Poly::from_coefficients */
static inline uint64_t reduce_i64(int64_t x, uint64_t q) {
  if (x >= 0) return (uint64_t)x % q;
  uint64_t m = (uint64_t)(-(x + 1)) + 1; /* |x| without overflow */
  uint64_t r = m % q;
  return r ? q - r : 0;
}
static uint32_t brv(uint32_t i, uint32_t bits) { uint32_t r = 0; for (uint32_t b = 0; b < bits; b++) { r = (r << 1) | (i & 1); i >>= 1; } return r; }
static uint32_t ilog2(uint32_t x) { uint32_t r = 0; while ((1u << r) < x) r++; return r; }

/* ------------------------------------------------------------------------------------------------ */
/* negacyclic NTT, natural-order in, bit-reversed-order out (fhe-math NttOperator)       */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { uint64_t w[PVWO_MAX_ELL], winv[PVWO_MAX_ELL], ninv; } ntt_tab;

static void ntt_tab_init(ntt_tab *t, uint64_t q, uint64_t psi, uint32_t ell) {
  uint32_t lg = ilog2(ell);
  uint64_t psi_inv = powmod(psi, q - 2, q);
  for (uint32_t i = 0; i < ell; i++) {
    t->w[i] = powmod(psi, brv(i, lg), q);             /* omegas[i] = psi^brv(i) */
    t->winv[i] = powmod(psi_inv, brv(i, lg), q);
  }
  t->ninv = powmod(ell, q - 2, q);
}
static void ntt_fwd(uint64_t *a, const ntt_tab *tb, uint64_t q, uint32_t ell) {
  const bmod bm = bmod_init(q);
  uint32_t t = ell;
  for (uint32_t m = 1; m < ell; m <<= 1) {
    t >>= 1;
    for (uint32_t i = 0; i < m; i++) {
      uint64_t s = tb->w[m + i];
      uint32_t j1 = 2 * i * t;
      for (uint32_t j = j1; j < j1 + t; j++) {
        uint64_t u = a[j], v = mulmod_b(a[j + t], s, &bm);
        a[j] = addmod(u, v, q);
        a[j + t] = submod(u, v, q);
      }
    }
  }
}
static void ntt_inv(uint64_t *a, const ntt_tab *tb, uint64_t q, uint32_t ell) {
  const bmod bm = bmod_init(q);
  uint32_t t = 1;
  for (uint32_t m = ell; m > 1; m >>= 1) {
    uint32_t h = m >> 1, j1 = 0;
    for (uint32_t i = 0; i < h; i++) {
      uint64_t s = tb->winv[h + i];
      for (uint32_t j = j1; j < j1 + t; j++) {
        uint64_t u = a[j], v = a[j + t];
        a[j] = addmod(u, v, q);
        a[j + t] = mulmod_b(submod(u, v, q), s, &bm);
      }
      j1 += 2 * t;
    }
    t <<= 1;
  }
  for (uint32_t j = 0; j < ell; j++) a[j] = mulmod_b(a[j], tb->ninv, &bm);
}

typedef struct { const pvwo_params *p; ntt_tab tab[PVWO_MAX_L]; bmod bm[PVWO_MAX_L]; } octx;
static void octx_init(octx *c, const pvwo_params *p) {
  c->p = p;
  for (uint32_t j = 0; j < p->L; j++) { ntt_tab_init(&c->tab[j], p->moduli[j], p->psi[j], p->ell); c->bm[j] = bmod_init(p->moduli[j]); }
}

/* small signed coefficients (ell) -> RNS -> NTT, out u64[L][ell]
 * (Poly::from_coefficients + change_representation(Ntt): encryption.rs:147-154, secret_key.rs:98-112,
 *  sample_error_1/2 parameters.rs:264-284) */
static void small_to_ntt(const octx *c, const int64_t *coef, uint64_t *out) {
  const pvwo_params *p = c->p;
  for (uint32_t j = 0; j < p->L; j++) {
    uint64_t q = p->moduli[j], *row = out + (size_t)j * p->ell;
    for (uint32_t t = 0; t < p->ell; t++) row[t] = reduce_i64(coef[t], q);
    ntt_fwd(row, &c->tab[j], q, p->ell);
  }
}
/* encode_scalar, parameters.rs:346-367: (m as i64) * [1, D, ..., D^(l-1)] -> RNS -> NTT */
static void encode_scalar(const octx *c, uint64_t m, uint64_t *out) {
  const pvwo_params *p = c->p;
  int64_t ms = (int64_t)m; /* `scalars[p] as i64`, encryption.rs:195 */
  for (uint32_t j = 0; j < p->L; j++) {
    uint64_t q = p->moduli[j], *row = out + (size_t)j * p->ell, mr = reduce_i64(ms, q);
    for (uint32_t t = 0; t < p->ell; t++) row[t] = mulmod_b(mr, p->gadget_rns[(size_t)j * p->ell + t], &c->bm[j]);
    ntt_fwd(row, &c->tab[j], q, p->ell);
  }
}

/* acc[L][ell] += a[L][ell] (.) b[L][ell]  -- Poly Mul then Poly Add, slot-wise (crs.rs:197-198) */
static inline void poly_mac(const octx *c, uint64_t *acc, const uint64_t *a, const uint64_t *b) {
  const pvwo_params *p = c->p;
  for (uint32_t j = 0; j < p->L; j++) {
    const bmod *bm = &c->bm[j];
    for (uint32_t t = 0; t < p->ell; t++) {
      size_t o = (size_t)j * p->ell + t;
      acc[o] = addmod(acc[o], mulmod_b(a[o], b[o], bm), bm->q);
    }
  }
}
static inline void poly_add(const pvwo_params *p, uint64_t *acc, const uint64_t *a) {
  for (uint32_t j = 0; j < p->L; j++)
    for (uint32_t t = 0; t < p->ell; t++) { size_t o = (size_t)j * p->ell + t; acc[o] = addmod(acc[o], a[o], p->moduli[j]); }
}
static inline void poly_sub(const pvwo_params *p, uint64_t *acc, const uint64_t *a) {
  for (uint32_t j = 0; j < p->L; j++)
    for (uint32_t t = 0; t < p->ell; t++) { size_t o = (size_t)j * p->ell + t; acc[o] = submod(acc[o], a[o], p->moduli[j]); }
}

/* ------------------------------------------------------------------------------------------------ */
/* multi-precision integers (num-bigint BigInt semantics: sign + magnitude, truncated division)      */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { int n; int neg; uint64_t w[BW]; } big;

static void big_norm(big *a) { while (a->n > 0 && a->w[a->n - 1] == 0) a->n--; if (a->n == 0) a->neg = 0; }
static void big_from_words(big *a, const uint64_t *w, int n) { memcpy(a->w, w, (size_t)n * 8); a->n = n; a->neg = 0; big_norm(a); }
static void big_from_u64(big *a, uint64_t v) { a->w[0] = v; a->n = 1; a->neg = 0; big_norm(a); }
static int mag_cmp(const big *a, const big *b) {
  if (a->n != b->n) return a->n < b->n ? -1 : 1;
  for (int i = a->n - 1; i >= 0; i--) if (a->w[i] != b->w[i]) return a->w[i] < b->w[i] ? -1 : 1;
  return 0;
}
static void mag_add(big *r, const big *a, const big *b) { /* r = |a| + |b| */
  int n = a->n > b->n ? a->n : b->n; u128 c = 0;
  for (int i = 0; i < n; i++) { c += (i < a->n ? a->w[i] : 0); c += (i < b->n ? b->w[i] : 0); r->w[i] = (uint64_t)c; c >>= 64; }
  r->w[n] = (uint64_t)c; r->n = n + 1;
}
static void mag_sub(big *r, const big *a, const big *b) { /* r = |a| - |b|, requires |a| >= |b| */
  uint64_t br = 0;
  for (int i = 0; i < a->n; i++) {
    uint64_t x = a->w[i], y = i < b->n ? b->w[i] : 0;
    uint64_t d = x - y, b1 = x < y; uint64_t d2 = d - br, b2 = d < br;
    r->w[i] = d2; br = b1 | b2;
  }
  r->n = a->n;
}
static void big_add_signed(big *r, const big *a, const big *b, int negate_b) {
  int bneg = b->n ? (b->neg ^ negate_b) : 0;
  big t;
  if (a->neg == bneg) { mag_add(&t, a, b); t.neg = a->neg; }
  else {
    int c = mag_cmp(a, b);
    if (c >= 0) { mag_sub(&t, a, b); t.neg = a->neg; } else { mag_sub(&t, b, a); t.neg = bneg; }
  }
  big_norm(&t); *r = t;
}
static void big_add(big *r, const big *a, const big *b) { big_add_signed(r, a, b, 0); }
static void big_sub(big *r, const big *a, const big *b) { big_add_signed(r, a, b, 1); }
static void big_mul(big *r, const big *a, const big *b) {
  big t; int n = a->n + b->n;
  memset(t.w, 0, (size_t)(n + 1) * 8);
  for (int i = 0; i < a->n; i++) {
    u128 c = 0;
    for (int j = 0; j < b->n; j++) { c += (u128)a->w[i] * b->w[j] + t.w[i + j]; t.w[i + j] = (uint64_t)c; c >>= 64; }
    t.w[i + b->n] = (uint64_t)c;
  }
  t.n = n; t.neg = (a->n && b->n) ? (a->neg ^ b->neg) : 0; big_norm(&t); *r = t;
}
/* magnitude division, Knuth vol.2 4.3.1 algorithm D with 64-bit digits: q = floor(|a|/|b|), r = |a| - q|b| */
static void mag_divrem(big *q, big *r, const big *a, const big *b) {
  if (mag_cmp(a, b) < 0) { big ra = *a; ra.neg = 0; q->n = 0; q->neg = 0; *r = ra; return; }
  if (b->n == 1) {
    uint64_t d = b->w[0]; u128 rem = 0; big qq; qq.n = a->n; qq.neg = 0;
    for (int i = a->n - 1; i >= 0; i--) { u128 cur = (rem << 64) | a->w[i]; qq.w[i] = (uint64_t)(cur / d); rem = cur % d; }
    big_norm(&qq); *q = qq; big_from_u64(r, (uint64_t)rem); return;
  }
  int n = b->n, m = a->n - b->n, s = __builtin_clzll(b->w[n - 1]);
  uint64_t u[BW + 1], v[BW];
  for (int i = n - 1; i > 0; i--) v[i] = s ? (b->w[i] << s) | (b->w[i - 1] >> (64 - s)) : b->w[i];
  v[0] = b->w[0] << s;
  u[a->n] = s ? a->w[a->n - 1] >> (64 - s) : 0;
  for (int i = a->n - 1; i > 0; i--) u[i] = s ? (a->w[i] << s) | (a->w[i - 1] >> (64 - s)) : a->w[i];
  u[0] = a->w[0] << s;
  big qq; qq.n = m + 1; qq.neg = 0;
  for (int j = m; j >= 0; j--) {
    u128 num = ((u128)u[j + n] << 64) | u[j + n - 1];
    u128 qhat = num / v[n - 1], rhat = num % v[n - 1];
    while (qhat >> 64 || (u128)(uint64_t)qhat * v[n - 2] > ((rhat << 64) | u[j + n - 2])) {
      qhat--; rhat += v[n - 1];
      if (rhat >> 64) break;
    }
    /* multiply and subtract */
    u128 borrow = 0, carry = 0;
    for (int i = 0; i < n; i++) {
      u128 pr = (u128)(uint64_t)qhat * v[i] + carry; carry = pr >> 64;
      uint64_t sub = (uint64_t)pr;
      u128 t = (u128)u[i + j] - sub - borrow;
      u[i + j] = (uint64_t)t; borrow = (t >> 64) & 1;
    }
    u128 t = (u128)u[j + n] - carry - borrow; u[j + n] = (uint64_t)t;
    if ((t >> 64) & 1) { /* add back */
      qhat--;
      u128 c = 0;
      for (int i = 0; i < n; i++) { c += (u128)u[i + j] + v[i]; u[i + j] = (uint64_t)c; c >>= 64; }
      u[j + n] += (uint64_t)c;
    }
    qq.w[j] = (uint64_t)qhat;
  }
  big_norm(&qq); *q = qq;
  big rr; rr.n = n; rr.neg = 0;
  for (int i = 0; i < n; i++) rr.w[i] = s ? (u[i] >> s) | (u[i + 1] << (64 - s)) : u[i];
  big_norm(&rr); *r = rr;
}
/* Rust BigInt `/` and `%`: truncated toward zero, remainder takes the dividend's sign */
static void big_tdivrem(big *q, big *r, const big *a, const big *b) {
  big qq, rr; mag_divrem(&qq, &rr, a, b);
  qq.neg = qq.n ? (a->neg ^ b->neg) : 0; rr.neg = rr.n ? a->neg : 0;
  if (q) *q = qq;
  if (r) *r = rr;
}
/* x mod Q into [0, Q) for signed x */
static void big_mod_floor(big *r, const big *x, const big *Q) {
  big rr; big_tdivrem(NULL, &rr, x, Q);
  if (rr.neg) big_add(&rr, &rr, Q);
  *r = rr;
}
static int big_cmp(const big *a, const big *b) { /* signed compare */
  if (a->neg != b->neg) return a->neg ? -1 : 1;
  int c = mag_cmp(a, b); return a->neg ? -c : c;
}
static void big_shr1(big *r, const big *a) { /* truncated |a|/2 with sign kept (BigInt / 2) */
  big t = *a;
  for (int i = 0; i < t.n; i++) t.w[i] = (t.w[i] >> 1) | (i + 1 < t.n ? t.w[i + 1] << 63 : 0);
  big_norm(&t); *r = t;
}
static void big_dbl(big *r, const big *a) { big_add(r, a, a); }

/* center_coefficient_with_precision, decryption.rs:139-152: x in [0,Q) -> x - Q if x > Q/2 */
static void centre(big *r, const big *x, const big *Q, const big *halfQ) {
  if (big_cmp(x, halfQ) > 0) big_sub(r, x, Q); else *r = *x;
}

/* CRT lift of one coefficient (RnsContext::lift): sum_j (x_j * inv_j mod q_j) * (Q/q_j) mod Q */
static void crt_lift(const pvwo_params *p, const uint64_t *res /* stride ell between limbs */, big *out, const big *Q) {
  big acc; acc.n = 0; acc.neg = 0;
  for (uint32_t j = 0; j < p->L; j++) {
    big y, qh, t;
    big_from_u64(&y, mulmod(res[(size_t)j * p->ell], p->qhat_inv[j], p->moduli[j]));
    big_from_words(&qh, p->qhat + (size_t)j * p->nw, (int)p->nw);
    big_mul(&t, &y, &qh);
    big_add(&acc, &acc, &t);
  }
  big_mod_floor(out, &acc, Q);
}

/* decode_scalar_pvw_rns, decryption.rs:10-58 + helpers :61-247, in the scalar form of SURVEY A.6:
 * every "constant polynomial" operation of the reference is arithmetic on one integer mod Q. */
static uint64_t decode(const octx *c, const uint64_t *zhat /* [L][ell] Ntt form */) {
  const pvwo_params *p = c->p;
  uint32_t ell = p->ell;
  uint64_t zc[PVWO_MAX_L * PVWO_MAX_ELL];
  memcpy(zc, zhat, (size_t)p->L * ell * 8);
  for (uint32_t j = 0; j < p->L; j++) ntt_inv(zc + (size_t)j * ell, &c->tab[j], p->moduli[j], ell);
  big Q, halfQ, D, M;
  big_from_words(&Q, p->Q, (int)p->nw); big_shr1(&halfQ, &Q);
  big_from_words(&D, p->delta, (int)p->nw);
  big_from_words(&M, p->delta_pow, (int)p->nw);
  big *z = (big *)malloc(sizeof(big) * ell), *tmp = (big *)malloc(sizeof(big) * ell);
  for (uint32_t t = 0; t < ell; t++) crt_lift(p, zc + t, &z[t], &Q);            /* Vec<BigUint>::from(&Poly), :118 */
  big a, b;
  for (uint32_t i = 0; i + 1 < ell; i++) {                                      /* :19-27 */
    big_mul(&a, &z[i], &D); big_sub(&a, &a, &z[i + 1]); big_mod_floor(&tmp[i], &a, &Q);
  }
  big last = tmp[0];                                                            /* :30-33 */
  for (uint32_t i = 1; i + 1 < ell; i++) { big_mul(&a, &last, &D); big_add(&a, &a, &tmp[i]); big_mod_floor(&last, &a, &Q); }
  /* reduce_modulo_poly :154-178 (mod_const = centre(D^(l-1) mod Q) = D^(l-1), always <= Q/2) */
  big pc, Mc, red, halfM, negHalfM;
  centre(&pc, &last, &Q, &halfQ);
  big_mod_floor(&a, &M, &Q); centre(&Mc, &a, &Q, &halfQ);
  big_tdivrem(NULL, &red, &pc, &Mc);
  big_shr1(&halfM, &Mc); negHalfM = halfM; if (negHalfM.n) negHalfM.neg ^= 1;
  if (big_cmp(&red, &halfM) > 0) big_sub(&red, &red, &Mc);
  else if (big_cmp(&red, &negHalfM) < 0) big_add(&red, &red, &Mc);
  big noise; big_mod_floor(&noise, &red, &Q);                                   /* bigints_to_poly of `reduced` */
  big Dc; big_mod_floor(&a, &D, &Q); centre(&Dc, &a, &Q, &halfQ);
  big twoD; big_dbl(&twoD, &Dc);
  for (int i = (int)ell - 2; i >= 0; i--) {                                     /* :44-48, divide_by_delta_rns :180-207 */
    big num, quo;
    big_sub(&a, &noise, &tmp[i]); big_mod_floor(&b, &a, &Q); centre(&num, &b, &Q, &halfQ);
    if (Dc.n == 0) { quo.n = 0; quo.neg = 0; }
    else {
      big_dbl(&a, &num);
      if (num.neg) big_sub(&a, &a, &Dc); else big_add(&a, &a, &Dc);
      big_tdivrem(&quo, NULL, &a, &twoD);
    }
    big_mod_floor(&noise, &quo, &Q);
  }
  /* plaintext = z0 * (-1) - noise_0, :51-53 ; extract_constant_term_as_u64 :226-247 */
  big pt; pt.n = 0; pt.neg = 0;
  big_sub(&a, &pt, &z[0]); big_sub(&a, &a, &noise); big_mod_floor(&b, &a, &Q); centre(&pt, &b, &Q, &halfQ);
  uint64_t result;
  if (pt.neg) {
    big thousand; big_from_u64(&thousand, 1000);
    big absv = pt; absv.neg = 0;
    if (big_cmp(&absv, &thousand) <= 0) result = 0;
    else {
      big_add(&a, &pt, &Q); big_tdivrem(NULL, &b, &a, &Q);
      result = (b.n <= 1 && !b.neg) ? (b.n ? b.w[0] : 0) : 0;                    /* to_u64().unwrap_or(0) */
    }
  } else result = (pt.n <= 1) ? (pt.n ? pt.w[0] : 0) : 0;
  free(z); free(tmp);
  return result;
}

/* ------------------------------------------------------------------------------------------------ */
/* exported entry points                                                                              */
/* ------------------------------------------------------------------------------------------------ */
int pvwo_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
/* launchers such as torch.distributed.run export OMP_NUM_THREADS=1: the timed CPU baseline sets its thread count explicitly */
void pvwo_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* coeffs i64[count][ell] -> out u64[count][L][ell] */
void pvwo_ntt_small(const pvwo_params *p, uint64_t count, const int64_t *coef, uint64_t *out) {
  octx c; octx_init(&c, p);
  size_t poly = (size_t)p->L * p->ell;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)count; i++) small_to_ntt(&c, coef + (size_t)i * p->ell, out + (size_t)i * poly);
}
/* in-place forward / inverse representation change of count polys u64[count][L][ell] */
void pvwo_ntt_poly(const pvwo_params *p, uint64_t count, uint64_t *polys, int inverse) {
  octx c; octx_init(&c, p);
  size_t poly = (size_t)p->L * p->ell;
  for (uint64_t i = 0; i < count; i++)
    for (uint32_t j = 0; j < p->L; j++) {
      uint64_t *row = polys + i * poly + (size_t)j * p->ell;
      if (inverse) ntt_inv(row, &c.tab[j], p->moduli[j], p->ell); else ntt_fwd(row, &c.tab[j], p->moduli[j], p->ell);
    }
}
void pvwo_encode_scalar(const pvwo_params *p, uint64_t m, uint64_t *out) { octx c; octx_init(&c, p); encode_scalar(&c, m, out); }

/* PublicKey::generate, public_key.rs:111-147 with explicit error; crs.rs:152-165 (transposed index):
 * b[cidx] = sum_j NTT(s[j]) * A[j][cidx] + NTT(e[cidx]).   A u64[k][k][L][ell]; sk, e i64[P][k][ell]; out u64[P][k][L][ell] */
void pvwo_keygen(const pvwo_params *p, uint64_t nparties, const uint64_t *A, const int64_t *sk, const int64_t *e, uint64_t *out) {
  octx c; octx_init(&c, p);
  size_t poly = (size_t)p->L * p->ell; uint32_t k = p->k;
#pragma omp parallel
  {
    uint64_t *shat = (uint64_t *)malloc(poly * k * 8), *tmp = (uint64_t *)malloc(poly * 8);
#pragma omp for schedule(dynamic)
    for (int64_t pi = 0; pi < (int64_t)nparties; pi++) {
      for (uint32_t j = 0; j < k; j++) small_to_ntt(&c, sk + ((size_t)pi * k + j) * p->ell, shat + (size_t)j * poly);
      for (uint32_t ci = 0; ci < k; ci++) {
        uint64_t *b = out + ((size_t)pi * k + ci) * poly;
        memset(b, 0, poly * 8);
        for (uint32_t j = 0; j < k; j++) poly_mac(&c, b, shat + (size_t)j * poly, A + ((size_t)j * k + ci) * poly);
        small_to_ntt(&c, e + ((size_t)pi * k + ci) * p->ell, tmp);
        poly_add(p, b, tmp);
      }
    }
    free(shat); free(tmp);
  }
}

/* encrypt with explicit randomness for D dealers (encryption.rs:105-214, 253-286).
 * A u64[k][k][L][ell]; B u64[nrows][k][L][ell] = rows [row0,row0+nrows) of the global key;
 * m u64[D][nrows]; r,e1 i64[D][k][ell]; e2 i64[D][nrows][ell];
 * c1 u64[D][k][L][ell] (NULL = skip), c2 u64[D][nrows][L][ell].  Parallel over dealers x parties. */
void pvwo_encrypt(const pvwo_params *p, uint64_t D, uint64_t nrows, const uint64_t *A, const uint64_t *B,
                  const uint64_t *m, const int64_t *r, const int64_t *e1, const int64_t *e2, uint64_t *c1, uint64_t *c2) {
  octx c; octx_init(&c, p);
  size_t poly = (size_t)p->L * p->ell; uint32_t k = p->k;
  uint64_t *rhat = (uint64_t *)malloc(poly * k * D * 8);
  pvwo_ntt_small(p, D * k, r, rhat);                                            /* :147-154 */
  if (c1) {
#pragma omp parallel
    {
      uint64_t *tmp = (uint64_t *)malloc(poly * 8);
#pragma omp for collapse(2) schedule(dynamic, 8)
      for (int64_t d = 0; d < (int64_t)D; d++)
        for (int64_t i = 0; i < (int64_t)k; i++) {                              /* crs.rs:187-199 */
          uint64_t *o = c1 + ((size_t)d * k + i) * poly;
          memset(o, 0, poly * 8);
          for (uint32_t j = 0; j < k; j++) poly_mac(&c, o, A + ((size_t)i * k + j) * poly, rhat + ((size_t)d * k + j) * poly);
          small_to_ntt(&c, e1 + ((size_t)d * k + i) * p->ell, tmp);             /* :161-173 */
          poly_add(p, o, tmp);
        }
      free(tmp);
    }
  }
  if (c2) {
#pragma omp parallel
    {
      uint64_t *tmp = (uint64_t *)malloc(poly * 8);
#pragma omp for collapse(2) schedule(dynamic, 8)
      for (int64_t d = 0; d < (int64_t)D; d++)
        for (int64_t pi = 0; pi < (int64_t)nrows; pi++) {                       /* :177-200 */
          uint64_t *o = c2 + ((size_t)d * nrows + pi) * poly;
          memset(o, 0, poly * 8);
          for (uint32_t j = 0; j < k; j++) poly_mac(&c, o, B + ((size_t)pi * k + j) * poly, rhat + ((size_t)d * k + j) * poly);
          encode_scalar(&c, m[(size_t)d * nrows + pi], tmp); poly_add(p, o, tmp);
          small_to_ntt(&c, e2 + ((size_t)d * nrows + pi) * p->ell, tmp); poly_add(p, o, tmp);
        }
      free(tmp);
    }
  }
  free(rhat);
}

/* zhat = sum_j NTT(s[j]) * c1[j] - c2p  (decryption.rs:257-274); out u64[L][ell] */
static void noisy_message(const octx *c, const uint64_t *shat, const uint64_t *c1, const uint64_t *c2p, uint64_t *z) {
  const pvwo_params *p = c->p; size_t poly = (size_t)p->L * p->ell;
  memset(z, 0, poly * 8);
  for (uint32_t j = 0; j < p->k; j++) poly_mac(c, z, shat + (size_t)j * poly, c1 + (size_t)j * poly);
  poly_sub(p, z, c2p);
}

/* decrypt_party_value for P parties x D dealers (decryption.rs:249-325).
 * sk i64[P][k][ell]; c1 u64[D][k][L][ell]; c2 u64[D][P][L][ell] (the c2 rows of exactly these P parties);
 * out u64[P][D]; zhat_out u64[P][D][L][ell] or NULL.  Like the reference, NTT(s) is formed per call, but
 * once per party rather than once per (ciphertext, j) (secret_key.rs:98-112) -- an optimistic baseline. */
void pvwo_decrypt(const pvwo_params *p, uint64_t P, uint64_t D, const int64_t *sk, const uint64_t *c1, const uint64_t *c2,
                  uint64_t *out, uint64_t *zhat_out) {
  octx c; octx_init(&c, p);
  size_t poly = (size_t)p->L * p->ell; uint32_t k = p->k;
  uint64_t *shat = (uint64_t *)malloc(poly * k * P * 8);
  pvwo_ntt_small(p, P * k, sk, shat);
#pragma omp parallel
  {
    uint64_t *z = (uint64_t *)malloc(poly * 8);
#pragma omp for collapse(2) schedule(dynamic, 4)
    for (int64_t pi = 0; pi < (int64_t)P; pi++)
      for (int64_t d = 0; d < (int64_t)D; d++) {
        noisy_message(&c, shat + (size_t)pi * k * poly, c1 + (size_t)d * k * poly, c2 + ((size_t)d * P + pi) * poly, z);
        if (zhat_out) memcpy(zhat_out + ((size_t)pi * D + d) * poly, z, poly * 8);
        out[(size_t)pi * D + d] = decode(&c, z);
      }
    free(z);
  }
  free(shat);
}

/* decode only: zhat u64[count][L][ell] -> out u64[count] */
void pvwo_decode(const pvwo_params *p, uint64_t count, const uint64_t *zhat, uint64_t *out) {
  octx c; octx_init(&c, p);
  size_t poly = (size_t)p->L * p->ell;
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t i = 0; i < (int64_t)count; i++) out[i] = decode(&c, zhat + (size_t)i * poly);
}

/* CRT lift of the power-basis coefficients of count polys: in u64[count][L][ell] (PowerBasis), out u64[count][ell][nw] */
void pvwo_lift(const pvwo_params *p, uint64_t count, const uint64_t *polys, uint64_t *out) {
  big Q; big_from_words(&Q, p->Q, (int)p->nw);
  size_t poly = (size_t)p->L * p->ell;
  for (uint64_t i = 0; i < count; i++)
    for (uint32_t t = 0; t < p->ell; t++) {
      big v; crt_lift(p, polys + i * poly + t, &v, &Q);
      uint64_t *o = out + (i * p->ell + t) * p->nw;
      memset(o, 0, (size_t)p->nw * 8); memcpy(o, v.w, (size_t)v.n * 8);
    }
}
