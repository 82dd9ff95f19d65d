/*
 * reheatfunq_b200 — host-side restatement of the d_min-restricted sample
 * enumeration (src/coverings/dminpermutations.cpp:90-879 and
 * the Cython glue reheatfunq/coverings/rdisks.pyx:87-414).  Combinatorial,
 * sequential, O(samples * N): it stays on the host, but it is what feeds the
 * bootstrap sample sets of HeatFlowAnomalyPosterior for dmin > 0
 * (posterior.py:299-343), so the drop-in needs it.
 *
 * No Boost.Geometry: the neighbour graph is built with the quadratic loop the
 * reference itself uses for N < 100 (dminpermutations.cpp:98-115); the rtree
 * branch only changes the complexity, not the result.
 *
 * Random numbers come from the caller's numpy BitGenerator through its public
 * bitgen_t interface, with numpy's documented masked-rejection
 * `random_interval`, so that a given Generator state yields the same samples
 * as the reference.
 */
#include <map>
#include <stack>
#include <cmath>
#include <limits>

namespace rhfq {
namespace coverings {

/* numpy/random/bitgen.h */
struct bitgen_t {
	void* state;
	uint64_t (*next_uint64)(void* st);
	uint32_t (*next_uint32)(void* st);
	double (*next_double)(void* st);
	uint64_t (*next_raw)(void* st);
};

/* numpy's random_interval: uniform integer in [0, max] by masked rejection */
inline uint64_t random_interval(bitgen_t* bg, uint64_t max)
{
	if (max == 0)
		return 0;
	uint64_t mask = max, value;
	mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4;
	mask |= mask >> 8; mask |= mask >> 16; mask |= mask >> 32;
	if (max <= 0xffffffffULL) {
		while ((value = (bg->next_uint32(bg->state) & mask)) > max) {}
	} else {
		while ((value = (bg->next_uint64(bg->state) & mask)) > max) {}
	}
	return value;
}

typedef std::vector<std::vector<size_t>> graph_t;

/* dminpermutations.cpp:90-157 */
inline size_t compute_neighbors(const double* xy, size_t N, double dmin, graph_t& nodes)
{
	nodes.assign(N, std::vector<size_t>());
	const double dmin2 = dmin * dmin;
	size_t links = 0;
	for (size_t i = 0; i < N; ++i) {
		const double xi = xy[2 * i], yi = xy[2 * i + 1];
		for (size_t j = i + 1; j < N; ++j) {
			const double dx = xi - xy[2 * j], dy = yi - xy[2 * j + 1];
			if (dx * dx + dy * dy <= dmin2) {
				nodes[i].push_back(j);
				nodes[j].push_back(i);
				++links;
			}
		}
	}
	return links;
}

/* dminpermutations.cpp:164-275 */
class NodeStack {
public:
	struct fork_t { size_t i; double ln_p; };
	explicit NodeStack(const graph_t& nodes) : blocked_(nodes.size(), 0), blocker_(nodes.size(), 0),
	                                           nodes_(nodes), n_blocked_(0) {}
	void clear() {
		stack_.clear();
		std::fill(blocked_.begin(), blocked_.end(), 0);
		n_blocked_ = 0;
	}
	bool empty() const { return stack_.empty(); }
	size_t top_node() const { return stack_.back().i; }
	double top_ln_p() const { return stack_.back().ln_p; }
	size_t blocked() const { return n_blocked_; }
	bool is_blocked(size_t i) const { return blocked_[i] != 0; }
	const std::vector<fork_t>& items() const { return stack_; }
	size_t free_index(size_t n = 0) const {
		size_t i = 0;
		++n;
		for (; i < nodes_.size(); ++i) {
			if (!blocked_[i]) {
				--n;
				if (n == 0)
					break;
			}
		}
		return i;
	}
	void push(size_t i) {
		const double ln_p0 = stack_.empty() ? 0.0 : stack_.back().ln_p;
		const double ln_p = ln_p0 - std::log((double)(nodes_.size() - n_blocked_));
		for (size_t n : nodes_[i]) {
			if (!blocked_[n]) { blocked_[n] = 1; blocker_[n] = i; ++n_blocked_; }
		}
		if (blocked_[i])
			throw std::runtime_error("blocking blocked!");
		blocked_[i] = 1; blocker_[i] = i; ++n_blocked_;
		stack_.push_back({i, ln_p});
	}
	void pop() {
		const size_t i = stack_.back().i;
		stack_.pop_back();
		for (size_t n : nodes_[i]) {
			if (blocked_[n] && blocker_[n] == i) { blocked_[n] = 0; --n_blocked_; }
		}
		blocked_[i] = 0;
		--n_blocked_;
	}
private:
	std::vector<fork_t> stack_;
	std::vector<char> blocked_;
	std::vector<size_t> blocker_;
	const graph_t& nodes_;
	size_t n_blocked_;
};

typedef std::map<std::vector<size_t>, double> sample_map_t;

/* dminpermutations.cpp:289-324 */
inline void add_sample(const NodeStack& stack, sample_map_t& samples)
{
	std::vector<size_t> sample;
	sample.reserve(stack.items().size());
	for (const auto& f : stack.items()) sample.push_back(f.i);
	std::sort(sample.begin(), sample.end());
	auto it = samples.lower_bound(sample);
	if (it == samples.end() || it->first != sample)
		samples.emplace_hint(it, std::move(sample), stack.top_ln_p());
	else
		it->second += std::log1p(std::exp(stack.top_ln_p() - it->second));
}

/* advance to the next permutation in lexicographic order (dminpermutations.cpp:395-423) */
inline void next_permutation(NodeStack& st, size_t n_nodes)
{
	while (!st.empty()) {
		size_t i = st.top_node();
		st.pop();
		bool success = false;
		for (++i; i < n_nodes; ++i) {
			if (!st.is_blocked(i)) { st.push(i); success = true; break; }
		}
		if (success)
			break;
	}
}

/* dminpermutations.cpp:359-433; returns false when max_samples / max_iter was exceeded */
inline bool dsr_deterministic(const graph_t& nodes, size_t max_samples, size_t max_iter,
                              sample_map_t& samples)
{
	NodeStack st(nodes);
	st.push(0);
	size_t iter = 0;
	while (samples.size() < max_samples && iter++ < max_iter && !st.empty()) {
		while (st.blocked() < nodes.size()) {
			const size_t i = st.free_index();
			if (i == nodes.size())
				throw std::runtime_error("Could not find free node although there should be.");
			st.push(i);
		}
		add_sample(st, samples);
		next_permutation(st, nodes.size());
	}
	return st.empty();
}

/* dminpermutations.cpp:446-513 */
inline void dsr_monte_carlo(const graph_t& nodes, bitgen_t* bg, size_t max_samples,
                            size_t max_iter, sample_map_t& samples)
{
	NodeStack st(nodes);
	for (size_t k = 0; k < max_iter; ++k) {
		if (samples.size() == max_samples)
			break;
		st.clear();
		if (k < nodes.size())
			st.push(k);
		else
			st.push((size_t)random_interval(bg, nodes.size() - 1));
		size_t free;
		while ((free = nodes.size() - st.blocked())) {
			const size_t i = st.free_index((size_t)random_interval(bg, free - 1));
			if (i == nodes.size())
				throw std::runtime_error("Could not find free node although there should be.");
			st.push(i);
		}
		add_sample(st, samples);
	}
}

struct RestrictedSamples {
	std::vector<double> w;
	std::vector<int64_t> off;
	std::vector<int64_t> idx;
};

/* dminpermutations.cpp:525-679 */
inline void determine_restricted_samples(const double* xy, size_t N, double dmin,
                                         size_t max_samples, size_t max_iter, bitgen_t* bg,
                                         RestrictedSamples& out)
{
	graph_t nodes;
	const size_t links = compute_neighbors(xy, N, dmin, nodes);
	out.w.clear(); out.off.assign(1, 0); out.idx.clear();
	if (links == 0) {
		out.w.push_back(1.0);
		for (size_t i = 0; i < N; ++i) out.idx.push_back((int64_t)i);
		out.off.push_back((int64_t)N);
		return;
	}
	/* nodes without neighbours are part of every sample */
	std::vector<size_t> always, n2i, i2n(N, (size_t)-1);
	graph_t packed;
	for (size_t i = 0; i < N; ++i) {
		if (nodes[i].empty()) always.push_back(i);
		else { i2n[i] = packed.size(); n2i.push_back(i); packed.push_back(nodes[i]); }
	}
	for (auto& nb : packed)
		for (size_t& k : nb) k = i2n[k];

	sample_map_t samples;
	if (!dsr_deterministic(packed, max_samples, max_iter, samples)) {
		if (!bg)
			throw std::runtime_error("More than the supplied upper limit of samples.");
		samples.clear();
		dsr_monte_carlo(packed, bg, max_samples, max_iter, samples);
	}
	/* normalise (dminpermutations.cpp:640-663) */
	long double max_log = -std::numeric_limits<long double>::infinity();
	for (const auto& s : samples) if (s.second > max_log) max_log = s.second;
	long double norm = 0.0;
	std::vector<double> p;
	p.reserve(samples.size());
	for (const auto& s : samples) { p.push_back((double)std::exp((long double)s.second - max_log)); norm += p.back(); }
	if (!(norm > 0.0))
		throw std::runtime_error("determine_restricted_samples: Norm is zero.");
	size_t k = 0;
	for (const auto& s : samples) {
		out.w.push_back((double)(p[k++] / norm));
		std::vector<size_t> full;
		for (size_t i : s.first) full.push_back(n2i[i]);
		full.insert(full.end(), always.begin(), always.end());
		std::sort(full.begin(), full.end());
		for (size_t i : full) out.idx.push_back((int64_t)i);
		out.off.push_back((int64_t)out.idx.size());
	}
}

/* dminpermutations.cpp:682-879 */
inline double global_PHmax(const double* xy, const double* PHmax_i, size_t N, double dmin)
{
	graph_t nodes;
	compute_neighbors(xy, N, dmin, nodes);
	struct component_t { std::vector<size_t> nodes; double PHmin, PHmax; };
	std::vector<component_t> comps;
	std::vector<char> handled(N, 0);
	for (size_t i = 0; i < N; ++i) {
		if (handled[i]) continue;
		comps.emplace_back();
		component_t& c = comps.back();
		std::stack<size_t> nb;
		nb.push(i);
		while (!nb.empty()) {
			const size_t j = nb.top();
			nb.pop();
			handled[j] = 1;
			c.nodes.push_back(j);
			for (size_t k : nodes[j]) {
				if (handled[k]) continue;
				nb.push(k);
				handled[k] = 1;
			}
		}
		std::sort(c.nodes.begin(), c.nodes.end());
		c.PHmin = std::numeric_limits<double>::infinity();
		c.PHmax = -std::numeric_limits<double>::infinity();
		for (size_t j : c.nodes) { c.PHmin = std::min(c.PHmin, PHmax_i[j]); c.PHmax = std::max(c.PHmax, PHmax_i[j]); }
	}
	std::sort(comps.begin(), comps.end(),
	          [](const component_t& l, const component_t& r) { return l.PHmin < r.PHmin; });
	{
		double min_PHmax = std::numeric_limits<double>::infinity();
		for (const auto& c : comps) min_PHmax = std::min(min_PHmax, c.PHmax);
		size_t keep = 0;
		while (keep < comps.size() && comps[keep].PHmin < min_PHmax) ++keep;
		comps.resize(std::max<size_t>(keep, 1));
	}
	std::vector<size_t> i2n(N, (size_t)-1), n2i(N, (size_t)-1);
	double PHmax_global = std::numeric_limits<double>::infinity();
	for (const component_t& c : comps) {
		double local;
		if (c.nodes.size() == 1) local = c.PHmin;
		else {
			graph_t ln;
			for (size_t i : c.nodes) { i2n[i] = ln.size(); n2i[ln.size()] = i; ln.push_back(nodes[i]); }
			for (auto& nb : ln) for (size_t& i : nb) i = i2n[i];
			NodeStack st(ln);
			st.push(0);
			local = -std::numeric_limits<double>::infinity();
			const unsigned long long max_iter = 10000000000ULL;
			unsigned long long iter = 0;
			while (iter++ < max_iter && !st.empty()) {
				while (st.blocked() < ln.size()) {
					const size_t i = st.free_index();
					if (i == ln.size())
						throw std::runtime_error("Could not find free node although there should be.");
					st.push(i);
				}
				double fork_min = std::numeric_limits<double>::infinity();
				for (const auto& f : st.items()) fork_min = std::min(fork_min, PHmax_i[n2i[f.i]]);
				local = std::max(local, fork_min);
				next_permutation(st, ln.size());
			}
			if (!st.empty()) local = std::nan("");
		}
		PHmax_global = std::min(PHmax_global, local);
	}
	return PHmax_global;
}

/* rdisks.pyx:87-148: one random conforming subset; mask[i] = 1 if kept */
inline void restricted_sample(const double* xy, size_t N, double dmin, bitgen_t* bg,
                              std::vector<char>& keep)
{
	std::vector<size_t> order;
	order.reserve(N);
	std::vector<char> mask(N, 1);
	for (size_t i = 0; i < N; ++i) {
		const size_t k = (i < N - 1) ? (size_t)random_interval(bg, N - i - 1) : 0;
		size_t l = 0;
		size_t j = mask[l] ? 1 : 0;
		while (j <= k) { ++l; if (mask[l]) ++j; }
		mask[l] = 0;
		order.push_back(l);
	}
	std::vector<char> removed(N, 0);
	const double dmin2 = dmin * dmin;
	for (size_t i = 0; i < N; ++i) {
		const size_t o0 = order[i];
		if (removed[o0]) continue;
		const double xi = xy[2 * o0], yi = xy[2 * o0 + 1];
		for (size_t j = i + 1; j < N; ++j) {
			const size_t o1 = order[j];
			if (removed[o1]) continue;
			const double dx = xi - xy[2 * o1], dy = yi - xy[2 * o1 + 1];
			if (dx * dx + dy * dy < dmin2) removed[o1] = 1;
		}
	}
	keep.assign(N, 0);
	for (size_t i = 0; i < N; ++i) keep[i] = !removed[i];
}

/* rdisks.pyx:229-318: B random conforming subsets, duplicates merged, weights = multiplicity / B */
inline void bootstrap_data_selection(const double* xy, size_t N, double dmin, size_t B,
                                     bitgen_t* bg, RestrictedSamples& out)
{
	std::map<std::vector<char>, size_t> seen;
	std::vector<std::vector<char>> order;
	std::vector<double> mult;
	std::vector<char> keep;
	for (size_t b = 0; b < B; ++b) {
		restricted_sample(xy, N, dmin, bg, keep);
		auto it = seen.find(keep);
		if (it == seen.end()) {
			seen.emplace(keep, order.size());
			order.push_back(keep);
			mult.push_back(1.0);
		} else {
			mult[it->second] += 1.0;
		}
	}
	out.w.clear(); out.off.assign(1, 0); out.idx.clear();
	for (size_t s = 0; s < order.size(); ++s) {
		out.w.push_back(mult[s] / (double)B);
		for (size_t i = 0; i < N; ++i) if (order[s][i]) out.idx.push_back((int64_t)i);
		out.off.push_back((int64_t)out.idx.size());
	}
}

} // namespace coverings
} // namespace rhfq
