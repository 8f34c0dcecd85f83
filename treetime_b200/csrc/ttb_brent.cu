// ttb_brent.cu -- lock-step Brent minimiser on the device (A8): one thread per branch runs the state machine of
// scipy.optimize.minimize_scalar(method='brent', bracket=(xa, xb, xc), tol=tol) -- the call the reference makes per
// branch in GTR.optimal_t_compressed (gtr.py:879-891) -- on the branch objective
//   cost(s) = -prob_t_profiles((pp, pc), multiplicity, s^2, return_log=True) + exp(s^4 / 10000)     (gtr.py:876-878)
// whose likelihood part is evaluated for ALL branches by one launch of branch_eval_kernel per iteration.  The state
// never leaves the device; between the evaluation and the update the caller may all-reduce the partial objective
// values over the pattern shards (ttb_brent_f_device_ptr), so every rank takes identical decisions.
// Per element the arithmetic is that of scipy's `Brent.optimize` (scipy/optimize/_optimize.py: _mintol = 1e-11,
// _cg = 0.3819660, convergence |x - xmid| < 2 tol1 - (b - a)/2), the same restatement as treetime_b200/brent.py.
// Compiled with -fmad=false: no fused multiply-adds, so that the decisions match the host restatement.
#include "ttb_brent.h"

namespace {

constexpr double MINTOL = 1.0e-11;
constexpr double CG = 0.3819660;

__device__ __forceinline__ double cost_of(double f, double s) { return -1.0 * f + exp(pow(s, 4.0) / 10000.0); }

// Convergence test + next trial point for one branch (the top half of scipy's loop body).
__device__ void propose(const TtbBrent& B, int i, double tol, int maxiter) {
  bool active = B.active[i] != 0;
  const double x = B.x[i], a = B.a[i], b = B.b[i];
  if (active && B.nit[i] >= maxiter) active = false;
  const double tol1 = tol * fabs(x) + MINTOL;
  const double tol2 = 2.0 * tol1;
  const double xmid = 0.5 * (a + b);
  if (active && fabs(x - xmid) < (tol2 - 0.5 * (b - a))) active = false;
  B.active[i] = active ? 1 : 0;
  if (!active) {
    B.ts[i] = -1.0;
    return;
  }
  const double w = B.w[i], v = B.v[i], fx = B.fx[i], fw = B.fw[i], fv = B.fv[i];
  double deltax = B.deltax[i], rat = B.rat[i];
  bool parabolic = false;
  if (!(fabs(deltax) <= tol1)) {
    double tmp1 = (x - w) * (fx - fv);
    double tmp2 = (x - v) * (fx - fw);
    double p = (x - v) * tmp2 - (x - w) * tmp1;
    tmp2 = 2.0 * (tmp2 - tmp1);
    if (tmp2 > 0.0) p = -p;
    tmp2 = fabs(tmp2);
    const double dx_temp = deltax;
    if ((p > tmp2 * (a - x)) && (p < tmp2 * (b - x)) && (fabs(p) < fabs(0.5 * tmp2 * dx_temp))) {
      parabolic = true;
      deltax = rat;                 // scipy: deltax = rat (the previous step), then the new rat
      rat = p / tmp2;
      const double u = x + rat;
      if ((u - a) < tol2 || (b - u) < tol2) rat = (xmid - x >= 0) ? tol1 : -tol1;
    }
  }
  if (!parabolic) {
    deltax = (x >= xmid) ? a - x : b - x;
    rat = CG * deltax;
  }
  const double u = (fabs(rat) < tol1) ? ((rat >= 0) ? x + tol1 : x - tol1) : x + rat;
  B.deltax[i] = deltax;
  B.rat[i] = rat;
  B.u[i] = u;
  B.ts[i] = u * u;
}

}  // namespace

// stage 0 / 1 / 2: the three bracket points have been evaluated in turn; stage >= 3: a trial point u.
__global__ void ttb_brent_step_kernel(TtbBrent B, int n, int stage, double tol, int maxiter) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (stage == -1) {                         // start: first bracket point
    B.ts[i] = B.xa[i] * B.xa[i];
    return;
  }
  if (stage == 0) {
    B.fa[i] = cost_of(B.f[i], B.xa[i]);
    B.ts[i] = B.xb[i] * B.xb[i];
    return;
  }
  if (stage == 1) {
    B.fb[i] = cost_of(B.f[i], B.xb[i]);
    B.ts[i] = B.xc[i] * B.xc[i];
    return;
  }
  if (stage == 2) {
    const double fc = cost_of(B.f[i], B.xc[i]);
    const double xa = B.xa[i], xb = B.xb[i], xc = B.xc[i], fa = B.fa[i], fb = B.fb[i];
    B.fc[i] = fc;
    if (!((xa < xb) && (xb < xc))) atomicOr(B.flags + 1, 1);
    if (!((fb < fa) && (fb < fc))) atomicOr(B.flags + 1, 2);
    B.x[i] = B.w[i] = B.v[i] = xb;
    B.fx[i] = B.fw[i] = B.fv[i] = fb;
    B.a[i] = xa;
    B.b[i] = xc;
    B.deltax[i] = 0.0;
    B.rat[i] = 0.0;
    B.nit[i] = 0;
    B.nfev[i] = 3;
    B.active[i] = 1;
    propose(B, i, tol, maxiter);
    if (B.active[i]) atomicAdd(B.flags, 1);
    return;
  }
  if (B.active[i]) {                          // the bottom half of scipy's loop body
    const double u = B.u[i];
    const double fu = cost_of(B.f[i], u);
    double x = B.x[i], w = B.w[i], v = B.v[i], fx = B.fx[i], fw = B.fw[i], fv = B.fv[i], a = B.a[i], b = B.b[i];
    B.nfev[i] += 1;
    if (fu > fx) {
      if (u < x) a = u; else b = u;
      if ((fu <= fw) || (w == x)) {
        v = w; w = u; fv = fw; fw = fu;
      } else if ((fu <= fv) || (v == x) || (v == w)) {
        v = u; fv = fu;
      }
    } else {
      if (u >= x) a = x; else b = x;
      v = w; w = x; x = u;
      fv = fw; fw = fx; fx = fu;
    }
    B.x[i] = x; B.w[i] = w; B.v[i] = v; B.fx[i] = fx; B.fw[i] = fw; B.fv[i] = fv; B.a[i] = a; B.b[i] = b;
    B.nit[i] += 1;
    propose(B, i, tol, maxiter);
    if (B.active[i]) atomicAdd(B.flags, 1);
  }
}

void ttb_brent_step(const TtbBrent& B, int n, int stage, double tol, int maxiter, cudaStream_t s) {
  ttb_brent_step_kernel<<<(n + 127) / 128, 128, 0, s>>>(B, n, stage, tol, maxiter);
}
