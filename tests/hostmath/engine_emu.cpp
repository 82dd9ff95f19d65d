/*
 * TEST-ONLY helper: the product's batch engine (rhfq_engine.inl) compiled with
 * RHFQ_EMU, i.e. every "kernel" runs as a host loop.  It exists to debug the
 * level-synchronous pipeline logic in the build container, which has no GPU.
 * It is never built into or loaded by the package (symbols are emu_rhfq_*).
 */
#define RHFQ_EMU 1
#include "../../reheatfunq_b200/csrc/rhfq_engine.inl"
#include "../../reheatfunq_b200/csrc/rhfq_resilience.inl"
#include "../../reheatfunq_b200/csrc/rhfq_gcp.inl"
#include "../../reheatfunq_b200/csrc/rhfq_coverings.inl"
#include "../../reheatfunq_b200/csrc/rhfq_blitest.inl"
#include "../../include/reheatfunq_b200.h"
#define RHFQ_API(name) emu_rhfq_##name
#include "../../reheatfunq_b200/csrc/rhfq_capi.inl"

/* --- interpolant evaluation: the three formulations side by side (unit test) --- */
extern "C" void emu_test_interp(const double* f, int level, const double* zff, int nz,
                                double* out_bary, double* out_interior, double* out_cheb,
                                double* out_fine, int* fine_bad)
{
	using namespace rhfq;
	const int Rmax = level;
	std::vector<double> tab;
	make_cheb_table(Rmax, tab);
	const int n = (8 << level) + 1;
	ChebTable T{tab.data(), tab.data() + n, tab.data() + 2 * n};
	/* Chebyshev coefficients through the product's own DCT body */
	PostState ps;
	ps.status = ST_OK; ps.level[0] = level; ps.level[1] = level;
	std::vector<double> bf(2 * n), bc(2 * n, 0.0);
	for (int i = 0; i < n; ++i) { bf[i] = f[i]; bf[n + i] = f[i]; }
	std::vector<double> tw;
	make_fine_twiddle(Rmax, tw);
	FineTwiddle W{tw.data(), FINE_SIGMA * (8 << Rmax)};
	launch_bli_coeff(2, BliCoeffBody{&ps, bf.data(), bc.data(), Rmax, W}, *(Stream*)nullptr, nullptr);
	/* fine theta grid through the product's FFT build body */
	std::vector<double> fine(2 * FineBuildBody::stride(Rmax));
	int bad[2] = {0, 0};
	launch_fine_build(2, FineBuildBody{&ps, bc.data(), Rmax, W, fine.data(), bad, 0, 0, 0, nullptr},
	                  *(Stream*)nullptr, nullptr);
	*fine_bad = bad[0];
	for (int i = 0; i < nz; ++i) {
		const double ff = zff[i], fb = 1.0 - zff[i];
		out_fine[i] = fine_eval(fine.data(), FINE_SIGMA * (n - 1), ff, fb);
		const double zv = 2.0 * ff - 1.0;
		out_bary[i] = bli_interpolate(zv, ff, fb, f, 1, n, T);
		out_interior[i] = bli_interpolate_interior(zv, f, 1, n, T.val);
		out_cheb[i] = cheb_eval(bc.data(), n - 1, ff, fb);
	}
}
