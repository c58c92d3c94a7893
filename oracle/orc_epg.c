/* ORACLE (test infrastructure) — EPG decay curves.  See decaes_oracle.h for status. */
#define _GNU_SOURCE
#include <math.h>
#include <string.h>
#include "decaes_oracle.h"

#define ORC_PI 3.14159265358979323846264338327950288
static const long double ORC_D2R_L = 3.14159265358979323846264338327950288L / 180.0L;

/* Base.Math.sind: exact range reduction in degrees, then sin/cos kernels on an
 * extended-precision radian argument.  Used at src/EPGdecaycurve.jl:940 (m0 = sind(alpha/2)).
 * Here the extended precision is x87 long double. */
double orc_sind(double x) {
  if (isnan(x) || isinf(x)) return NAN;
  double rx = copysign(fmod(x, 360.0), x);
  double arx = fabs(rx);
  if (rx == 0.0) return rx;
  if (arx < 45.0) return (double)sinl((long double)rx * ORC_D2R_L);
  if (arx <= 135.0) return copysign((double)cosl((90.0L - (long double)arx) * ORC_D2R_L), rx);
  if (arx == 180.0) return copysign(0.0, rx);
  if (arx < 225.0) {
    long double y = (180.0L - (long double)arx) * (rx < 0 ? -1.0L : 1.0L);
    return (double)sinl(y * ORC_D2R_L);
  }
  if (arx <= 315.0) return -copysign((double)cosl((270.0L - (long double)arx) * ORC_D2R_L), rx);
  return (double)sinl(((long double)rx - (long double)copysign(360.0, rx)) * ORC_D2R_L);
}

static double orc_cosd(double x) { /* cosd(x) = sind(90 - x) up to the reduction; only used for d/dalpha */
  return (double)cosl((long double)x * ORC_D2R_L);
}

/* ------------------------------------------------------------------------------------------
 * epg_impulse_response! for EPGWork_ReIm_DualFlat_Split_Dynamic   src/EPGdecaycurve.jl:948-1028
 * followed by dc[i] = abs(sind(alpha/2) * dc[i])                  src/EPGdecaycurve.jl:936-946
 * 1-based indexing of the reference is kept through the W()/R() macros.
 * ---------------------------------------------------------------------------------------- */
void orc_epg_decay_curve(int ETL, double alpha_deg, double TE, double T2, double T1, double *dc,
                         double *work) {
  double *Wp = work, *Rp = work + 3 * (size_t)ETL; /* MPSV1 (written), MPSV2 (read) */
  const int dy = ETL, dz = 2 * ETL;
#define W(k) Wp[(k)-1]
#define R(k) Rp[(k)-1]
#define SWAP()      \
  do {              \
    double *t_ = Wp; \
    Wp = Rp;        \
    Rp = t_;        \
  } while (0)
  const double alpha = alpha_deg * (ORC_PI / 180.0); /* deg2rad  :955 */
  const double E1 = exp(-(TE / 2) / T1), E2 = exp(-(TE / 2) / T2); /* :958 */
  double sina, cosa;
  sincos(alpha, &sina, &cosa);
  const double E2h = (E2 * E2) / 2, E1E2 = E1 * E2, E1sq = E1 * E1; /* :960 */
  const double a = E2h, b = E2h * cosa, c = E1E2 * sina, d = E1sq * cosa; /* :961 */
  const double cp = -c / 2;                                               /* :962 */
  double F, Fb, Z, C, S, Cp, Sp;

  dc[0] = a - b; /* :966 */
  W(1) = a - b, W(1 + dy) = 0.0, W(1 + dz) = cp;
  W(2) = a + b, W(2 + dy) = 0.0, W(2 + dz) = 0.0;
  SWAP();

  for (int i = 2; i <= ETL / 2; i++) { /* :971-1000 */
    F = R(1), Fb = R(1 + dy), Z = R(1 + dz);
    C = F + Fb, S = F - Fb;
    Cp = a * C, Sp = b * S;
    dc[i - 1] = W(1) = fma(-c, Z, Cp - Sp);
    W(2) = fma(c, Z, Cp + Sp);
    W(1 + dz) = fma(cp, S, d * Z);
    for (int k = 2; k <= i; k++) { /* :979-993 (loop body and the k == i tail are identical) */
      F = R(k), Fb = R(k + dy), Z = R(k + dz);
      C = F + Fb, S = F - Fb;
      Cp = a * C, Sp = b * S;
      W(k + 1) = fma(c, Z, Cp + Sp);
      W(k - 1 + dy) = fma(-c, Z, Cp - Sp);
      W(k + dz) = fma(cp, S, d * Z);
    }
    W(i + dy) = 0.0; /* :995-997 */
    W(i + 1 + dy) = 0.0;
    W(i + 1 + dz) = 0.0;
    SWAP();
  }

  for (int i = ETL / 2 + 1; i <= ETL - 1; i++) { /* :1002-1020 */
    F = R(1), Fb = R(1 + dy), Z = R(1 + dz);
    C = F + Fb, S = F - Fb;
    Cp = a * C, Sp = b * S;
    dc[i - 1] = W(1) = fma(-c, Z, Cp - Sp);
    W(2) = fma(c, Z, Cp + Sp);
    W(1 + dz) = fma(cp, S, d * Z);
    for (int k = 2; k <= ETL - i + 1; k++) {
      F = R(k), Fb = R(k + dy), Z = R(k + dz);
      C = F + Fb, S = F - Fb;
      Cp = a * C, Sp = b * S;
      W(k + 1) = fma(c, Z, Cp + Sp);
      W(k - 1 + dy) = fma(-c, Z, Cp - Sp);
      W(k + dz) = fma(cp, S, d * Z);
    }
    SWAP();
  }

  F = R(1), Fb = R(1 + dy), Z = R(1 + dz); /* :1022-1024 */
  C = F + Fb, S = F - Fb;
  dc[ETL - 1] = fma(-c, Z, fma(a, C, (-b) * S));

  const double m0 = orc_sind(alpha_deg / 2); /* :940-943 */
  for (int i = 0; i < ETL; i++) dc[i] = fabs(m0 * dc[i]);
#undef W
#undef R
}

/* ------------------------------------------------------------------------------------------
 * Value and derivative with respect to alpha (per DEGREE) of the curve above.
 * The reference obtains this with ForwardDiff.jacobian! through the very same code
 * (EPGJacobianFunctor, src/EPGdecaycurve.jl:224-248; used at src/T2mapSEcorr.jl:299-308,
 * 362-371).  This is the same forward-mode differentiation written out by hand: every
 * quantity carries (value, d/dalpha); abs follows ForwardDiff (sign flip of the partial when
 * signbit(value)); sind' = (pi/180) cosd; deg2rad' = pi/180.
 * work: 12*ETL doubles.
 * ---------------------------------------------------------------------------------------- */
void orc_epg_decay_curve_jac(int ETL, double alpha_deg, double TE, double T2, double T1,
                             double *dc, double *ddc, double *work) {
  double *Wp = work, *Rp = work + 3 * (size_t)ETL;
  double *dWp = work + 6 * (size_t)ETL, *dRp = work + 9 * (size_t)ETL;
  const int dy = ETL, dz = 2 * ETL;
#define W(k) Wp[(k)-1]
#define R(k) Rp[(k)-1]
#define DW(k) dWp[(k)-1]
#define DR(k) dRp[(k)-1]
#define SWAP2()      \
  do {               \
    double *t_ = Wp;  \
    Wp = Rp;         \
    Rp = t_;         \
    t_ = dWp;        \
    dWp = dRp;       \
    dRp = t_;        \
  } while (0)
  const double kk = ORC_PI / 180.0;
  const double alpha = alpha_deg * kk;
  const double E1 = exp(-(TE / 2) / T1), E2 = exp(-(TE / 2) / T2);
  double sina, cosa;
  sincos(alpha, &sina, &cosa);
  const double dsina = cosa * kk, dcosa = -sina * kk;
  const double E2h = (E2 * E2) / 2, E1E2 = E1 * E2, E1sq = E1 * E1;
  const double a = E2h, b = E2h * cosa, c = E1E2 * sina, d = E1sq * cosa, cp = -c / 2;
  const double db = E2h * dcosa, dc_ = E1E2 * dsina, dd = E1sq * dcosa, dcp = -dc_ / 2;
  double F, Fb, Z, C, S, Cp, Sp, dF, dFb, dZ, dC, dS, dCp, dSp;

#define LOAD(k)                                   \
  F = R(k), Fb = R((k) + dy), Z = R((k) + dz);    \
  dF = DR(k), dFb = DR((k) + dy), dZ = DR((k) + dz); \
  C = F + Fb, S = F - Fb, dC = dF + dFb, dS = dF - dFb; \
  Cp = a * C, Sp = b * S, dCp = a * dC, dSp = db * S + b * dS
#define V_FB() fma(-c, Z, Cp - Sp)
#define D_FB() ((-dc_) * Z + (-c) * dZ + (dCp - dSp))
#define V_F() fma(c, Z, Cp + Sp)
#define D_F() (dc_ * Z + c * dZ + (dCp + dSp))
#define V_Z() fma(cp, S, d * Z)
#define D_Z() (dcp * S + cp * dS + (dd * Z + d * dZ))

  double *ir = dc, *dir = ddc; /* impulse response and its derivative, scaled at the end */
  ir[0] = a - b, dir[0] = -db;
  W(1) = a - b, W(1 + dy) = 0.0, W(1 + dz) = cp;
  DW(1) = -db, DW(1 + dy) = 0.0, DW(1 + dz) = dcp;
  W(2) = a + b, W(2 + dy) = 0.0, W(2 + dz) = 0.0;
  DW(2) = db, DW(2 + dy) = 0.0, DW(2 + dz) = 0.0;
  SWAP2();

  for (int i = 2; i <= ETL - 1; i++) {
    const int first_half = (i <= ETL / 2);
    LOAD(1);
    ir[i - 1] = W(1) = V_FB(), dir[i - 1] = DW(1) = D_FB();
    W(2) = V_F(), DW(2) = D_F();
    W(1 + dz) = V_Z(), DW(1 + dz) = D_Z();
    const int kmax = first_half ? i : ETL - i + 1;
    for (int k = 2; k <= kmax; k++) {
      LOAD(k);
      W(k + 1) = V_F(), DW(k + 1) = D_F();
      W(k - 1 + dy) = V_FB(), DW(k - 1 + dy) = D_FB();
      W(k + dz) = V_Z(), DW(k + dz) = D_Z();
    }
    if (first_half) {
      W(i + dy) = 0.0, W(i + 1 + dy) = 0.0, W(i + 1 + dz) = 0.0;
      DW(i + dy) = 0.0, DW(i + 1 + dy) = 0.0, DW(i + 1 + dz) = 0.0;
    }
    SWAP2();
  }
  F = R(1), Fb = R(1 + dy), Z = R(1 + dz);
  dF = DR(1), dFb = DR(1 + dy), dZ = DR(1 + dz);
  C = F + Fb, S = F - Fb, dC = dF + dFb, dS = dF - dFb;
  ir[ETL - 1] = fma(-c, Z, fma(a, C, (-b) * S));
  dir[ETL - 1] = (-dc_) * Z + (-c) * dZ + (a * dC - (db * S + b * dS));

  const double m0 = orc_sind(alpha_deg / 2);
  const double dm0 = orc_cosd(alpha_deg / 2) * kk / 2; /* d/dalpha sind(alpha/2) */
  for (int i = 0; i < ETL; i++) {
    double v = m0 * ir[i];
    double dv = dm0 * ir[i] + m0 * dir[i];
    dc[i] = fabs(v);
    ddc[i] = signbit(v) ? -dv : dv;
  }
#undef W
#undef R
#undef DW
#undef DR
#undef LOAD
}

/* ------------------------------------------------------------------------------------------
 * epg_decay_curve! for EPGWork_ReIm_DualVector_Split_Dynamic with EPGOptions (beta != 180)
 * src/EPGdecaycurve.jl:722-818.  MPSV entries are 3-vectors (F, Fbar, Z) stored contiguously.
 * work: 6*ETL doubles.
 * ---------------------------------------------------------------------------------------- */
static inline double dot3(const double *u, const double *v) {
  return u[0] * v[0] + u[1] * v[1] + u[2] * v[2];
}

void orc_epg_decay_curve_beta(int ETL, double alpha_deg, double TE, double T2, double T1,
                              double beta_deg, double *dc, double *work) {
  double *M1 = work, *M2 = work + 3 * (size_t)ETL;
#define V1(j) (M1 + 3 * ((j)-1))
#define V2(j) (M2 + 3 * ((j)-1))
#define SET(p, x, y, z) ((p)[0] = (x), (p)[1] = (y), (p)[2] = (z))
#define SWAPM()     \
  do {              \
    double *t_ = M1; \
    M1 = M2;        \
    M2 = t_;        \
  } while (0)
  const double A = alpha_deg / 180; /* B1correction :34 */
  const double a1r = (A * 180) * (ORC_PI / 180.0);
  const double air = (A * beta_deg) * (ORC_PI / 180.0);
  const double E1 = exp(-(TE / 2) / T1), E2 = exp(-(TE / 2) / T2);
  double sh, ch;
  sincos(a1r / 2, &sh, &ch);
  const double s2h = sh * sh, c2h = ch * ch;
  const double sin1 = 2 * sh * ch;
  double sini, cosi;
  sincos(air, &sini, &cosi);
  const double c2hi = (1 + cosi) / 2, s2hi = 1 - c2hi;
  const double E2sq = E2 * E2;
  const double a1 = E2sq * c2h, b1 = E2sq * s2h, c1 = E1 * E2 * sin1;
  const double ai = E2sq * c2hi, bi = E2sq * s2hi, ci = E1 * E2 * sini, di = E1 * E1 * cosi;
  const double Fv[3] = {ai, bi, ci}, Fbv[3] = {bi, ai, -ci}, Zv[3] = {-ci / 2, ci / 2, di};
  double FM0, FbM0, ZM0, FM1, FbM1, ZM1, FM2, FbM2, ZM2;

  /* i = 1 */
  const double m0 = sh;
  SET(V1(1), b1 * m0, 0.0, -c1 * m0 / 2);
  SET(V1(2), a1 * m0, 0.0, 0.0);
  dc[0] = fabs(V1(1)[0]);
  SWAPM();
  /* i = 2 */
  FM0 = dot3(Fv, V2(1)), FbM0 = dot3(Fbv, V2(1)), ZM0 = dot3(Zv, V2(1));
  FM1 = dot3(Fv, V2(2)), FbM1 = dot3(Fbv, V2(2)), ZM1 = dot3(Zv, V2(2));
  SET(V1(1), FbM0, FbM1, ZM0);
  SET(V1(2), FM0, 0.0, ZM1);
  SET(V1(3), FM1, 0.0, 0.0);
  dc[1] = fabs(FbM0);
  SWAPM();

  for (int i = 3; i <= ETL - 1; i++) {
    const int first_half = (i <= ETL / 2);
    FM0 = dot3(Fv, V2(1)), FbM0 = dot3(Fbv, V2(1)), ZM0 = dot3(Zv, V2(1));
    FM1 = dot3(Fv, V2(2)), FbM1 = dot3(Fbv, V2(2)), ZM1 = dot3(Zv, V2(2));
    FM2 = dot3(Fv, V2(3)), FbM2 = dot3(Fbv, V2(3)), ZM2 = dot3(Zv, V2(3));
    SET(V1(1), FbM0, FbM1, ZM0);
    SET(V1(2), FM0, FbM2, ZM1);
    const int jmax = first_half ? i - 1 : ETL - i;
    for (int j = 3; j <= jmax; j++) {
      FM0 = FM1, FM1 = FM2, ZM1 = ZM2;
      FM2 = dot3(Fv, V2(j + 1)), FbM2 = dot3(Fbv, V2(j + 1)), ZM2 = dot3(Zv, V2(j + 1));
      SET(V1(j), FM0, FbM2, ZM1);
    }
    if (first_half) {
      SET(V1(i), FM1, 0.0, ZM2);
      SET(V1(i + 1), FM2, 0.0, 0.0);
    }
    dc[i - 1] = fabs(FbM0);
    SWAPM();
  }
  dc[ETL - 1] = fabs(dot3(Fbv, V2(1)));
#undef V1
#undef V2
#undef SET
}

/* ------------------------------------------------------------------------------------------
 * Value + d/dalpha (per degree) of the general-beta curve: hand forward mode through the same
 * recursion (the reference differentiates it with ForwardDiff, src/EPGdecaycurve.jl:224-248).
 * In-place state update: old state M_j yields F.M_j -> new F_{j+1}, Fbar.M_j -> new Fbar_{j-1},
 * Z.M_j -> new Z_j, which is the double-buffered update of :722-818 with the buffers merged
 * (values are bitwise those of orc_epg_decay_curve_beta; tests/test_oracle_epg.py).
 * work: 12*ETL doubles.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  double v, d;
} dual_t;
static inline dual_t ddot3(const double *u, const double *du, dual_t x, dual_t y, dual_t z) {
  dual_t r;
  r.v = u[0] * x.v + u[1] * y.v + u[2] * z.v;
  r.d = (du[0] * x.v + du[1] * y.v + du[2] * z.v) + (u[0] * x.d + u[1] * y.d + u[2] * z.d);
  return r;
}
void orc_epg_decay_curve_beta_jac(int ETL, double alpha_deg, double TE, double T2, double T1,
                                  double beta_deg, double *dc, double *ddc, double *work) {
  const double kk = ORC_PI / 180.0;
  const double A = alpha_deg / 180;
  const double a1r = (A * 180) * kk, air = (A * beta_deg) * kk;
  const double kb = (beta_deg / 180) * kk;
  const double E1 = exp(-(TE / 2) / T1), E2 = exp(-(TE / 2) / T2);
  double sh, ch, sini, cosi;
  sincos(a1r / 2, &sh, &ch);
  sincos(air, &sini, &cosi);
  const double dsh = ch * (kk / 2), dch = -sh * (kk / 2);
  const double s2h = sh * sh, c2h = ch * ch, ds2h = 2 * sh * dsh, dc2h = 2 * ch * dch;
  const double sin1 = 2 * sh * ch, dsin1 = 2 * (dsh * ch + sh * dch);
  const double dsini = cosi * kb, dcosi = -sini * kb;
  const double c2hi = (1 + cosi) / 2, s2hi = 1 - c2hi, dc2hi = dcosi / 2, ds2hi = -dc2hi;
  const double E2sq = E2 * E2, E1E2 = E1 * E2, E1sq = E1 * E1;
  const double a1 = E2sq * c2h, b1 = E2sq * s2h, c1 = E1E2 * sin1;
  const double da1 = E2sq * dc2h, db1 = E2sq * ds2h, dc1 = E1E2 * dsin1;
  const double ai = E2sq * c2hi, bi = E2sq * s2hi, ci = E1E2 * sini, di = E1sq * cosi;
  const double dai = E2sq * dc2hi, dbi = E2sq * ds2hi, dci = E1E2 * dsini, ddi = E1sq * dcosi;
  const double Fv[3] = {ai, bi, ci}, Fbv[3] = {bi, ai, -ci}, Zv[3] = {-ci / 2, ci / 2, di};
  const double dFv[3] = {dai, dbi, dci}, dFbv[3] = {dbi, dai, -dci}, dZv[3] = {-dci / 2, dci / 2, ddi};
  dual_t *F = (dual_t *)work, *Fb = F + ETL, *Z = Fb + ETL; /* 1-based use, ETL/2 + 2 entries needed */
  const dual_t zero = {0.0, 0.0};
  const double m0 = sh, dm0 = dsh;
#define EMIT(i, x) (dc[i] = fabs((x).v), ddc[i] = signbit((x).v) ? -(x).d : (x).d)
  F[1].v = b1 * m0, F[1].d = db1 * m0 + b1 * dm0, Fb[1] = zero;
  Z[1].v = -c1 * m0 / 2, Z[1].d = -(dc1 * m0 + c1 * dm0) / 2;
  F[2].v = a1 * m0, F[2].d = da1 * m0 + a1 * dm0, Fb[2] = zero, Z[2] = zero;
  EMIT(0, F[1]);
  for (int i = 2; i <= ETL - 1; i++) {
    const int first_half = (i <= ETL / 2);
    const int nproc = first_half ? i : ETL - i + 1;
    dual_t FM = ddot3(Fv, dFv, F[1], Fb[1], Z[1]), FbM = ddot3(Fbv, dFbv, F[1], Fb[1], Z[1]),
           ZM = ddot3(Zv, dZv, F[1], Fb[1], Z[1]);
    EMIT(i - 1, FbM);
    F[1] = FbM, Z[1] = ZM;
    dual_t pend = FM;
    for (int j = 2; j <= nproc; j++) {
      FM = ddot3(Fv, dFv, F[j], Fb[j], Z[j]), FbM = ddot3(Fbv, dFbv, F[j], Fb[j], Z[j]),
      ZM = ddot3(Zv, dZv, F[j], Fb[j], Z[j]);
      F[j] = pend;
      pend = FM;
      Fb[j - 1] = FbM;
      Z[j] = ZM;
    }
    if (first_half) F[nproc + 1] = pend, Fb[nproc] = zero, Fb[nproc + 1] = zero, Z[nproc + 1] = zero;
  }
  {
    dual_t last = ddot3(Fbv, dFbv, F[1], Fb[1], Z[1]);
    EMIT(ETL - 1, last);
  }
#undef EMIT
}
