/*
 * TEST-ONLY helper: compiles the product's __host__ __device__ math header
 * (reheatfunq_b200/csrc/rhfq_math.cuh) with g++ so that the fixed-node
 * numerics can be checked against the oracle on a machine without a GPU.
 * Never loaded by the package; the product runs this math only inside the
 * sm_100a kernels.
 */
#include "../../reheatfunq_b200/csrc/rhfq_math.cuh"
#include <vector>
#include <cstring>

using namespace rhfq;

extern "C" {

int hm_sizeof_locals() { return (int)sizeof(SetLocals); }

/* Fills L (opaque buffer of hm_sizeof_locals() bytes) and k[N].  Returns status. */
int hm_locals(const double* q, const double* c, int N, double p, double s, double n,
              double v, double amin, void* Lbuf, double* k)
{
	SetLocals L;
	std::memset(&L, 0, sizeof(L));
	int st = compute_locals(q, c, N, p, s, n, v, amin, k, L);
	L.log_eps = std::log(DBL_EPS); L.sqrt_eps = SQRT_EPS;      /* the reference's double build */
	std::memcpy(Lbuf, &L, sizeof(L));
	return st;
}

int hm_log_scale(void* Lbuf, const double* k, double* log_scale)
{
	SetLocals* L = (SetLocals*)Lbuf;
	int st = log_integrand_max(*L, k, *log_scale);
	L->log_scale = *log_scale;
	return st;
}

int hm_ymax(void* Lbuf, double* ymax)
{
	SetLocals* L = (SetLocals*)Lbuf;
	int st = y_taylor_transition(*L, *ymax);
	L->ymax = *ymax;
	L->ztrans = 1.0 - *ymax;
	return st;
}

/* log I(z) unscaled; info[4*i..] = a_lo, a_mode, a_hi, evals */
int hm_log_outer(const void* Lbuf, const double* k, const double* z, int nz, double* out,
                 double* info)
{
	const SetLocals* L = (const SetLocals*)Lbuf;
	int worst = 0;
	for (int i = 0; i < nz; ++i) {
		QuadInfo qi;
		const double lsum = sum_log1p(k, L->N, z[i]);
		const double l1p_wz = log1p(-L->w * z[i]);
		int st = log_outer_unscaled(*L, lsum, l1p_wz, out[i], &qi);
		if (st > worst) worst = st;
		if (info) {
			info[4 * i] = qi.a_lo; info[4 * i + 1] = qi.a_mode;
			info[4 * i + 2] = qi.a_hi; info[4 * i + 3] = qi.evals;
		}
	}
	return worst;
}

int hm_log_large_z(const void* Lbuf, const double* y, int ny, int y_integrated, double* out,
                   double* info)
{
	const SetLocals* L = (const SetLocals*)Lbuf;
	int worst = 0;
	for (int i = 0; i < ny; ++i) {
		QuadInfo qi;
		int st = log_large_z_unscaled(*L, y[i], y_integrated != 0, out[i], &qi);
		if (st > worst) worst = st;
		if (info) {
			info[4 * i] = qi.a_lo; info[4 * i + 1] = qi.a_mode;
			info[4 * i + 2] = qi.a_hi; info[4 * i + 3] = qi.evals;
		}
	}
	return worst;
}

double hm_lgdiff(double a, double n, double v, double C, double lv) { return lgdiff(a, n, v, C, lv); }
double hm_digdiff(double a, double n, double v) { return digdiff(a, n, v); }
double hm_trigdiff(double a, double n, double v) { return trigdiff(a, n, v); }
/* lgamma-difference table of (n, v): size, build, and the three reads at a in [1, 1024) */
int hm_gtab_doubles() { return GTAB_STRIDE; }
void hm_gtab_build(double n, double v, double* tab)
{
	for (int j = 0; j < GTAB_SIZE; ++j) gtab_fill_record(n, v, j, tab + GTAB_REC * j);
}
void hm_gtab_eval(const double* tab, const double* a, int na, double* out)
{
	for (int i = 0; i < na; ++i) {
		out[3 * i] = gtab_eval_h(tab, a[i]);
		out[3 * i + 1] = gtab_eval_d1(tab, a[i]);
		out[3 * i + 2] = gtab_eval_d2(tab, a[i]);
	}
}
double hm_tetragamma(double x) { return tetragamma_pos(x); }

} // extern "C"

/* --- GCP internals (debug / unit tests) --- */
#define RHFQ_MATH_ONLY 1
#include <cstdint>
#include "../../reheatfunq_b200/csrc/rhfq_gcp.inl"
extern "C" {
double hm_gcp_ln_F(double a, double lp, double s, double n, double v)
{
	GcpTuple t{lp, s, log(s), n, v};
	return gcp_ln_F(a, t);
}
double hm_gcp_derivative_amax(double v, double n) { return gcp_derivative_amax(v, n); }
double hm_gcp_f0(double a, double v, double n) { return gcp_f0(a, v, n); }
int hm_gcp_integrals(double lp, double s, double n, double v, double amin, double* out)
{
	GcpTuple t{lp, s, log(s), n, v};
	GcpIntegrals I = gcp_integrals(t, amin, true);
	out[0] = I.lnPhi; out[1] = I.m1; out[2] = I.m2; out[3] = I.m3;
	return I.status;
}
}
