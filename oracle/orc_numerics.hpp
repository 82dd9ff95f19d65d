/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked or imported by the product
 * (reheatfunq_b200/); only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it.
 *
 * Numerical building blocks that the reference takes from Boost.Math 1.86
 * (not vendored under /root/reference, absent from this image).  Everything
 * here is restated from the published algorithms:
 *   - tanh-sinh / exp-sinh double-exponential quadrature (Takahasi & Mori 1974)
 *     with level doubling until |I_k - I_{k-1}| <= tol * L1, the termination
 *     contract the reference's call sites rely on
 *     (include/anomaly/posterior/integrand.hpp:414-437,
 *      include/anomaly/posterior/localsandnorm.hpp:221-226,
 *      include/anomaly/posterior/large_z.hpp:476-490,
 *      external/pdtoolbox/src/gamma_conjugate_prior.cpp:497-524,611-614)
 *   - Brent's minimiser (Brent 1973, ch. 5) as called with bits=1e6, 200 its
 *     (integrand.hpp:274-279, amax.hpp:114-120, large_z.hpp:282-287)
 *   - Algorithm 748 (Alefeld, Potra, Shi 1995) with Boost's eps_tolerance
 *     stopping rule (posterior.hpp:297-306, large_z.hpp:338-348)
 *   - digamma / trigamma / polygamma(2) (integrand.hpp:178,197;
 *     gamma_conjugate_prior.cpp:123-148,608)
 * Node-for-node identity with Boost is NOT claimed; results agree to the
 * quadrature tolerance ("parity pinning").
 */
#ifndef ORC_NUMERICS_HPP
#define ORC_NUMERICS_HPP

#include <cmath>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <utility>
#include <algorithm>
#include <string>

namespace orc {

template<typename R>
struct lim {
	static constexpr R eps() { return std::numeric_limits<R>::epsilon(); }
	static R root_eps() { return std::sqrt(std::numeric_limits<R>::epsilon()); }
	static constexpr R inf() { return std::numeric_limits<R>::infinity(); }
};

template<typename R> inline R pi_v() { return static_cast<R>(3.141592653589793238462643383279502884L); }

/* ------------------------------------------------------------------ *
 *  Kahan summation (include/numerics/kahan.hpp:99-125)                *
 * ------------------------------------------------------------------ */
#pragma GCC push_options
#pragma GCC optimize("-fno-associative-math")
template<typename R>
struct Kahan {
	R S = 0, corr = 0;
	void add(R x) {
		R dS = x - corr;
		R next = S + dS;
		corr = (next - S) - dS;
		S = next;
	}
	operator R() const { return S; }
};
#pragma GCC pop_options

/* ------------------------------------------------------------------ *
 *  Polygamma functions for x > 0: upward recurrence + asymptotic      *
 *  series (Abramowitz & Stegun 6.3.18, 6.4.12).                       *
 * ------------------------------------------------------------------ */
template<typename R>
R digamma(R x)
{
	if (!(x > 0))
		return std::numeric_limits<R>::quiet_NaN();
	R res = 0;
	while (x < 30) {
		res -= 1 / x;
		x += 1;
	}
	const R xi = 1 / x, x2 = xi * xi;
	/* B_2k / (2k): 1/12, -1/120, 1/252, -1/240, 1/132, -691/32760, 1/12, -3617/8160 */
	R ser = x2 * (R(1) / 12 - x2 * (R(1) / 120 - x2 * (R(1) / 252 - x2 * (R(1) / 240
	        - x2 * (R(1) / 132 - x2 * (R(691) / 32760 - x2 * (R(1) / 12 - x2 * R(3617) / 8160)))))));
	return res + std::log(x) - xi / 2 - ser;
}

template<typename R>
R trigamma(R x)
{
	if (!(x > 0))
		return std::numeric_limits<R>::quiet_NaN();
	R res = 0;
	while (x < 30) {
		res += 1 / (x * x);
		x += 1;
	}
	const R xi = 1 / x, x2 = xi * xi;
	/* B_2k: 1/6, -1/30, 1/42, -1/30, 5/66, -691/2730, 7/6, -3617/510 */
	R ser = xi * x2 * (R(1) / 6 - x2 * (R(1) / 30 - x2 * (R(1) / 42 - x2 * (R(1) / 30
	        - x2 * (R(5) / 66 - x2 * (R(691) / 2730 - x2 * (R(7) / 6 - x2 * R(3617) / 510)))))));
	return res + xi + x2 / 2 + ser;
}

/* psi''(x) = polygamma(2, x) */
template<typename R>
R tetragamma(R x)
{
	if (!(x > 0))
		return std::numeric_limits<R>::quiet_NaN();
	R res = 0;
	while (x < 30) {
		res -= 2 / (x * x * x);
		x += 1;
	}
	const R xi = 1 / x, x2 = xi * xi;
	/* -1/x^2 - 1/x^3 - sum B_2k (2k+1) / x^(2k+2) */
	R ser = x2 * x2 * (R(1) / 2 - x2 * (R(1) / 6 - x2 * (R(1) / 6 - x2 * (R(3) / 10
	        - x2 * (R(5) / 6 - x2 * (R(691) / 210 - x2 * (R(35) / 2)))))));
	return res - x2 - xi * x2 - ser;
}

/* ------------------------------------------------------------------ *
 *  Brent's minimiser on [lo, hi].  Starts from the upper end like the *
 *  reference's dependency does; `bits` is capped at digits/2.         *
 * ------------------------------------------------------------------ */
template<typename R, typename F>
std::pair<R, R> brent_minimize(F f, R lo, R hi, int bits, std::uintmax_t& max_iter)
{
	bits = std::min<int>(std::numeric_limits<R>::digits / 2, bits);
	const R tolerance = std::ldexp(R(1), 1 - bits);
	const R golden = R(0.3819660L);
	R x, w, v, u, fx, fw, fv, fu;
	R delta = 0, delta2 = 0;
	x = w = v = hi;
	fx = fw = fv = f(x);
	std::uintmax_t count = max_iter;
	do {
		const R mid = (lo + hi) / 2;
		const R fract1 = tolerance * std::fabs(x) + tolerance / 4;
		const R fract2 = 2 * fract1;
		if (std::fabs(x - mid) <= (fract2 - (hi - lo) / 2))
			break;
		bool golden_step = true;
		if (std::fabs(delta2) > fract1) {
			/* Parabolic fit through (v,w,x): */
			R r = (x - w) * (fx - fv);
			R q = (x - v) * (fx - fw);
			R p = (x - v) * q - (x - w) * r;
			q = 2 * (q - r);
			if (q > 0)
				p = -p;
			q = std::fabs(q);
			const R td = delta2;
			delta2 = delta;
			if (std::isfinite(p) && std::isfinite(q)
			    && !(std::fabs(p) >= std::fabs(q * td / 2))
			    && !(p <= q * (lo - x)) && !(p >= q * (hi - x)))
			{
				delta = p / q;
				u = x + delta;
				if ((u - lo) < fract2 || (hi - u) < fract2)
					delta = (mid - x) < 0 ? -std::fabs(fract1) : std::fabs(fract1);
				golden_step = false;
			}
		}
		if (golden_step) {
			delta2 = (x >= mid) ? lo - x : hi - x;
			delta = golden * delta2;
		}
		u = (std::fabs(delta) >= fract1) ? x + delta
		    : (delta > 0 ? x + std::fabs(fract1) : x - std::fabs(fract1));
		fu = f(u);
		if (fu <= fx) {
			if (u >= x) lo = x; else hi = x;
			v = w; w = x; x = u;
			fv = fw; fw = fx; fx = fu;
		} else {
			if (u < x) lo = u; else hi = u;
			if (fu <= fw || w == x) {
				v = w; w = u; fv = fw; fw = fu;
			} else if (fu <= fv || v == x || v == w) {
				v = u; fv = fu;
			}
		}
	} while (--count);
	max_iter -= count;
	return std::make_pair(x, fx);
}

/* ------------------------------------------------------------------ *
 *  Algorithm 748 (Alefeld, Potra, Shi).                               *
 * ------------------------------------------------------------------ */
template<typename R>
struct eps_tolerance {
	R eps;
	explicit eps_tolerance(unsigned bits) {
		eps = std::max<R>(std::ldexp(R(1), 1 - (int)bits), 4 * lim<R>::eps());
	}
	bool operator()(R a, R b) const {
		return std::fabs(a - b) <= eps * std::min(std::fabs(a), std::fabs(b));
	}
};

namespace detail748 {

template<typename R, typename F>
void bracket(F& f, R& a, R& b, R c, R& fa, R& fb, R& d, R& fd)
{
	const R tol = lim<R>::eps() * 2;
	if ((b - a) < 2 * tol * a)
		c = a + (b - a) / 2;
	else if (c <= a + std::fabs(a) * tol)
		c = a + std::fabs(a) * tol;
	else if (c >= b - std::fabs(b) * tol)
		c = b - std::fabs(b) * tol;
	R fc = f(c);
	if (fc == 0) {
		a = c; fa = 0; d = 0; fd = 0;
		return;
	}
	if ((fa < 0) != (fc < 0) && fa != 0) {
		d = b; fd = fb; b = c; fb = fc;
	} else {
		d = a; fd = fa; a = c; fa = fc;
	}
}

template<typename R>
R safe_div(R num, R denom, R r)
{
	if (std::fabs(denom) < 1) {
		if (std::fabs(denom * std::numeric_limits<R>::max()) <= std::fabs(num))
			return r;
	}
	return num / denom;
}

template<typename R>
R secant_interpolate(R a, R b, R fa, R fb)
{
	const R tol = lim<R>::eps() * 5;
	R c = a - (fa / (fb - fa)) * (b - a);
	if (c <= a + std::fabs(a) * tol || c >= b - std::fabs(b) * tol)
		return (a + b) / 2;
	return c;
}

template<typename R>
R quadratic_interpolate(R a, R b, R d, R fa, R fb, R fd, unsigned count)
{
	R B = safe_div<R>(fb - fa, b - a, std::numeric_limits<R>::max());
	R A = safe_div<R>(fd - fb, d - b, std::numeric_limits<R>::max());
	A = safe_div<R>(A - B, d - a, R(0));
	if (A == 0)
		return secant_interpolate(a, b, fa, fb);
	R c = ((A < 0) == (fa < 0)) ? a : b;
	for (unsigned i = 1; i <= count; ++i) {
		c -= safe_div<R>(fa + (B + A * (c - b)) * (c - a),
		                 B + A * (2 * c - a - b), 1 + c - a);
	}
	if (c <= a || c >= b)
		c = secant_interpolate(a, b, fa, fb);
	return c;
}

template<typename R>
R cubic_interpolate(R a, R b, R d, R e, R fa, R fb, R fd, R fe)
{
	/* Inverse cubic interpolation through the four points: */
	R q11 = (d - e) * fd / (fe - fd);
	R q21 = (b - d) * fb / (fd - fb);
	R q31 = (a - b) * fa / (fb - fa);
	R d21 = (b - d) * fd / (fd - fb);
	R d31 = (a - b) * fb / (fb - fa);
	R q22 = (d21 - q11) * fb / (fe - fb);
	R q32 = (d31 - q21) * fa / (fd - fa);
	R d32 = (d31 - q21) * fd / (fd - fa);
	R q33 = (d32 - q22) * fa / (fe - fa);
	R c = q31 + q32 + q33 + a;
	if (!(c > a) || !(c < b))
		c = quadratic_interpolate(a, b, d, fa, fb, fd, 3);
	return c;
}

} // namespace detail748

template<typename R, typename F, typename Tol>
std::pair<R, R> toms748(F f, R ax, R bx, R fax, R fbx, Tol tol, std::uintmax_t& max_iter)
{
	using namespace detail748;
	std::uintmax_t count = max_iter;
	if (count == 0)
		return std::make_pair(ax, bx);
	R a = ax, b = bx, fa = fax, fb = fbx;
	if (!(a < b))
		throw std::domain_error("toms748: a >= b");
	if (tol(a, b) || fa == 0 || fb == 0) {
		max_iter = 0;
		if (fa == 0) b = a;
		else if (fb == 0) a = b;
		return std::make_pair(a, b);
	}
	if ((fa < 0) == (fb < 0))
		throw std::domain_error("toms748: root not bracketed");
	const R mu = R(0.5);
	R d = 0, e = 0, fd = 0, fe = 0, c;
	fe = e = fd = R(1e5);
	if (fa != 0) {
		c = secant_interpolate(a, b, fa, fb);
		bracket(f, a, b, c, fa, fb, d, fd);
		--count;
		if (count && fa != 0 && !tol(a, b)) {
			c = quadratic_interpolate(a, b, d, fa, fb, fd, 2);
			e = d; fe = fd;
			bracket(f, a, b, c, fa, fb, d, fd);
			--count;
		}
	}
	while (count && fa != 0 && !tol(a, b)) {
		const R a0 = a, b0 = b;
		const R min_diff = std::numeric_limits<R>::min() * 32;
		bool prof = (std::fabs(fa - fb) < min_diff) || (std::fabs(fa - fd) < min_diff)
		         || (std::fabs(fa - fe) < min_diff) || (std::fabs(fb - fd) < min_diff)
		         || (std::fabs(fb - fe) < min_diff) || (std::fabs(fd - fe) < min_diff);
		if (prof)
			c = quadratic_interpolate(a, b, d, fa, fb, fd, 2);
		else
			c = cubic_interpolate(a, b, d, e, fa, fb, fd, fe);
		e = d; fe = fd;
		bracket(f, a, b, c, fa, fb, d, fd);
		if (--count == 0 || fa == 0 || tol(a, b))
			break;
		prof = (std::fabs(fa - fb) < min_diff) || (std::fabs(fa - fd) < min_diff)
		    || (std::fabs(fa - fe) < min_diff) || (std::fabs(fb - fd) < min_diff)
		    || (std::fabs(fb - fe) < min_diff) || (std::fabs(fd - fe) < min_diff);
		if (prof)
			c = quadratic_interpolate(a, b, d, fa, fb, fd, 3);
		else
			c = cubic_interpolate(a, b, d, e, fa, fb, fd, fe);
		bracket(f, a, b, c, fa, fb, d, fd);
		if (--count == 0 || fa == 0 || tol(a, b))
			break;
		/* Double-length secant step: */
		R u, fu;
		if (std::fabs(fa) < std::fabs(fb)) { u = a; fu = fa; }
		else { u = b; fu = fb; }
		c = u - 2 * (fu / (fb - fa)) * (b - a);
		if (std::fabs(c - u) > (b - a) / 2)
			c = a + (b - a) / 2;
		e = d; fe = fd;
		bracket(f, a, b, c, fa, fb, d, fd);
		if (--count == 0 || fa == 0 || tol(a, b))
			break;
		/* Additional bisection if convergence was slow: */
		if ((b - a) < mu * (b0 - a0))
			continue;
		e = d; fe = fd;
		bracket(f, a, b, R(a + (b - a) / 2), fa, fb, d, fd);
		--count;
	}
	max_iter -= count;
	if (fa == 0) b = a;
	else if (fb == 0) a = b;
	return std::make_pair(a, b);
}

template<typename R, typename F, typename Tol>
std::pair<R, R> toms748(F f, R ax, R bx, Tol tol, std::uintmax_t& max_iter)
{
	if (max_iter <= 2)
		throw std::runtime_error("toms748: need more than two iterations.");
	max_iter -= 2;
	R fa = f(ax), fb = f(bx);
	auto r = toms748<R>(f, ax, bx, fa, fb, tol, max_iter);
	max_iter += 2;
	return r;
}

/* ------------------------------------------------------------------ *
 *  Double-exponential quadrature.                                     *
 * ------------------------------------------------------------------ */
struct QuadratureError : public std::runtime_error {
	explicit QuadratureError(const std::string& s) : std::runtime_error(s) {}
};

/*
 * tanh-sinh on the finite interval [a,b].  f is called as f(x) with x computed
 * from the nearer end point so that no precision is lost close to the ends.
 */
template<typename R, typename F>
R tanh_sinh_finite(F f, R a, R b, R tol, R* error, R* L1, size_t* levels,
                   int min_level = 3, int max_level = 12)
{
	const R hpi = pi_v<R>() / 2;
	const R hw = (b - a) / 2;
	/* t beyond which 1 - |x| underflows the smallest normal number: */
	const R tmax = std::asinh(std::log(2 / std::numeric_limits<R>::min()) / (2 * hpi));
	const R eps = lim<R>::eps();

	auto node = [&](R t, R& xc, R& w) {
		/* xc = 1 - x = 2 / (1 + exp(2u)),  w = (pi/2) cosh t / cosh^2 u */
		const R u = hpi * std::sinh(t);
		const R e2u = std::exp(2 * u);
		xc = 2 / (1 + e2u);
		const R ch = std::cosh(u);
		w = hpi * std::cosh(t) / (ch * ch);
	};

	R fmax = 0;
	auto feval = [&](R x) -> R {
		R y = f(x);
		if (std::isnan(y))
			throw QuadratureError("tanh_sinh: NaN integrand.");
		if (std::isinf(y))
			throw QuadratureError("tanh_sinh: infinite integrand.");
		fmax = std::max(fmax, std::fabs(y));
		return y;
	};

	/* Level 0 with h = 1: centre + symmetric pairs */
	R h = 1;
	R f0 = feval(a + hw);
	R sum = hpi * f0;          /* weight at t=0 is pi/2 */
	R l1 = hpi * std::fabs(f0);
	auto add_pairs = [&](R t0, R dt) {
		/* add nodes t0, t0+dt, ... on both sides until negligible */
		int small = 0;
		for (R t = t0; t <= tmax; t += dt) {
			R xc, w;
			node(t, xc, w);
			if (w == 0 || xc == 0)
				break;
			const R xr = b - hw * xc;
			const R xl = a + hw * xc;
			R fr = (xr < b) ? feval(xr) : R(0);
			R fl = (xl > a) ? feval(xl) : R(0);
			sum += w * (fr + fl);
			l1 += w * (std::fabs(fr) + std::fabs(fl));
			if (w * fmax < R(1e-3) * eps * l1) {
				if (++small >= 2)
					break;
			} else {
				small = 0;
			}
		}
	};
	add_pairs(h, h);
	R I_prev = sum * h;
	R I = I_prev, err = lim<R>::inf();
	int k = 0;
	for (k = 1; k <= max_level; ++k) {
		h /= 2;
		add_pairs(h, 2 * h);
		I = sum * h;
		err = std::fabs(I - I_prev);
		I_prev = I;
		if (k >= min_level && err <= tol * (l1 * h))
			break;
	}
	if (error) *error = err * hw;
	if (L1) *L1 = l1 * h * hw;
	if (levels) *levels = (size_t)k;
	return I * hw;
}

/*
 * tanh-sinh on the half-infinite interval [a, inf) via x = a - 1 + 2/(1+t),
 * t in (-1,1)  (the substitution the reference's dependency documents).
 */
template<typename R, typename F>
R tanh_sinh_halfinf(F f, R a, R tol, R* error, R* L1, size_t* levels)
{
	auto g = [&](R t_from_minus1) -> R {
		/* s = 1 + t in (0,2): x = a - 1 + 2/s */
		const R s = t_from_minus1;
		if (s <= 0)
			return R(0);
		const R z = 1 / s;
		const R x = a - 1 + 2 * z;
		if (std::isinf(x))
			return R(0);
		const R fx = f(x);
		if (fx == 0)
			return R(0);
		return fx * z * z * 2;
	};
	/* Integrate g over s in (0,2): */
	return tanh_sinh_finite<R>(g, R(0), R(2), tol, error, L1, levels);
}

/*
 * exp-sinh on [a, inf): x = a + exp(pi/2 sinh t).
 */
template<typename R, typename F>
R exp_sinh(F f, R a, R tol, R* error, R* L1, size_t* levels,
           int min_level = 3, int max_level = 12)
{
	const R hpi = pi_v<R>() / 2;
	const R eps = lim<R>::eps();
	const R lmax = std::log(std::numeric_limits<R>::max()) - 8;
	const R tmax = std::asinh(lmax / hpi);

	R fmax = 0;
	auto term = [&](R t, R& absterm) -> R {
		const R u = hpi * std::sinh(t);
		const R ex = std::exp(u);
		const R w = hpi * std::cosh(t) * ex;
		if (w == 0 || ex == 0) { absterm = 0; return R(0); }
		const R x = a + ex;
		R y = f(x);
		if (std::isnan(y))
			throw QuadratureError("exp_sinh: NaN integrand.");
		if (std::isinf(y))
			throw QuadratureError("exp_sinh: infinite integrand.");
		fmax = std::max(fmax, std::fabs(y));
		absterm = std::fabs(y) * w;
		return y * w;
	};

	R h = 1;
	R sum = 0, l1 = 0;
	{
		R at;
		sum = term(0, at);
		l1 = at;
	}
	auto add_side = [&](R t0, R dt, int sgn) {
		int small = 0;
		R last = lim<R>::inf();
		for (R t = t0; t <= tmax; t += dt) {
			R at;
			R v = term(sgn * t, at);
			sum += v;
			l1 += at;
			/* The integrands on this path are bounded and unimodal: once the terms
			 * are decreasing (the mode has been passed) stop after three
			 * consecutive negligible ones.  Towards x -> a the weights decay
			 * double exponentially, which bounds the remainder by w * fmax. */
			const bool negligible = (sgn > 0)
			    ? (at <= last && l1 > 0 && at < R(1e-3) * eps * l1)
			    : (l1 > 0 && hpi * std::cosh(t) * std::exp(-hpi * std::sinh(t)) * fmax
			                 < R(1e-3) * eps * l1);
			last = at;
			if (negligible) {
				if (++small >= 3)
					break;
			} else {
				small = 0;
			}
		}
	};
	add_side(h, h, +1);
	add_side(h, h, -1);
	R I_prev = sum * h, I = I_prev, err = lim<R>::inf();
	int k;
	for (k = 1; k <= max_level; ++k) {
		h /= 2;
		add_side(h, 2 * h, +1);
		add_side(h, 2 * h, -1);
		I = sum * h;
		err = std::fabs(I - I_prev);
		I_prev = I;
		if (k >= min_level && err <= tol * (l1 * h))
			break;
	}
	if (error) *error = err;
	if (L1) *L1 = l1 * h;
	if (levels) *levels = (size_t)k;
	return I;
}

} // namespace orc

#endif
