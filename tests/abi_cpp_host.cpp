// TEST PROGRAM: a C++ host that uses libpfa through the C ABI only (include/pfa.h) - no Python, no torch - the way the PolyFEM
// shim does (polyfem_b200/host/assembler_shim.hpp). One Kuhn cell (6 P1 tets), NeoHookean:
//   argv[1] == "partition": host-only part (no GPU needed): pfa_partition_create for 2 ranks, invariants printed;
//   argv[1] == "assemble":  pfa_create + pfa_grad_hess with host pointers on device 0, then the same mesh as a 2-rank
//                           owner-computes partition (both handles on device 0): owned columns of the two ranks together
//                           reproduce the single-handle matrix, energies add up. Exit code 0 = all checks passed.
#include <pfa.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

static const int kTets[6][4] = {{0, 1, 3, 7}, {0, 1, 5, 7}, {0, 2, 3, 7}, {0, 2, 6, 7}, {0, 4, 5, 7}, {0, 4, 6, 7}};

int main(int argc, char **argv)
{
	const bool assemble = argc > 1 && std::strcmp(argv[1], "assemble") == 0;
	// 8 cube corners, corner c at (c & 1, (c >> 1) & 1, (c >> 2) & 1); 6 tets around the diagonal 0 - 7, positively oriented
	std::vector<double> xyz(8 * 3);
	for (int c = 0; c < 8; ++c)
		for (int d = 0; d < 3; ++d)
			xyz[c * 3 + d] = (c >> d) & 1;
	std::vector<int32_t> conn(6 * 4);
	std::vector<double> vert(6 * 12);
	for (int e = 0; e < 6; ++e)
	{
		int t[4] = {kTets[e][0], kTets[e][1], kTets[e][2], kTets[e][3]};
		auto det = [&](const int *q) {
			double a[3][3];
			for (int r = 0; r < 3; ++r)
				for (int d = 0; d < 3; ++d)
					a[r][d] = xyz[q[r + 1] * 3 + d] - xyz[q[0] * 3 + d];
			return a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) + a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
		};
		if (det(t) < 0)
			std::swap(t[1], t[2]);
		for (int k = 0; k < 4; ++k)
		{
			conn[e * 4 + k] = t[k];
			for (int d = 0; d < 3; ++d)
				vert[(e * 4 + k) * 3 + d] = xyz[t[k] * 3 + d];
		}
	}
	// ---- partition (host only) ----
	int owned_total = 0;
	std::vector<pfa_partition *> parts(2, nullptr);
	for (int r = 0; r < 2; ++r)
	{
		if (pfa_partition_create(6, 4, 8, conn.data(), 2, r, &parts[r]) != PFA_OK)
			return std::printf("pfa_partition_create failed\n"), 2;
		int32_t n_own, n_ghost, n_local, n_owned;
		pfa_partition_sizes(parts[r], &n_own, &n_ghost, &n_local, &n_owned);
		std::printf("rank %d: %d own + %d ghost elements, %d local nodes, %d owned\n", r, n_own, n_ghost, n_local, n_owned);
		owned_total += n_owned;
	}
	if (owned_total != 8)
		return std::printf("every node must be owned exactly once\n"), 3;
	// ---- sparsity pattern without a device: in one Kuhn cell the diagonal 0 - 7 touches every node, the other pairs as the tets list them
	{
		pfa_host_pattern *hp = nullptr;
		if (pfa_host_pattern_create(6, 4, 8, conn.data(), &hp) != PFA_OK)
			return std::printf("pfa_host_pattern_create failed\n"), 10;
		int64_t n_pairs = 0;
		const int32_t *adj_off, *adj, *slot;
		pfa_host_pattern_arrays(hp, &n_pairs, &adj_off, &adj, &slot);
		bool ok = adj_off[1] - adj_off[0] == 8 && adj_off[8] - adj_off[7] == 8 && adj_off[8] == n_pairs;
		for (int e = 0; e < 6 && ok; ++e)
			for (int i = 0; i < 4; ++i)
				for (int j = 0; j < 4; ++j)
					ok = ok && adj[slot[(e * 4 + i) * 4 + j]] == conn[e * 4 + i];
		std::printf("host pattern: %lld node pairs\n", (long long)n_pairs);
		pfa_host_pattern_destroy(hp);
		if (!ok)
			return std::printf("host pattern: wrong adjacency or slot map\n"), 11;
	}
	if (!assemble)
	{
		for (auto *p : parts)
			pfa_partition_destroy(p);
		return 0;
	}
	// ---- single handle ----
	const double w[1] = {1.0 / 6.0};
	const double rg[12] = {-1, -1, -1, 1, 0, 0, 0, 1, 0, 0, 0, 1}; // P1 reference gradients (auto_p_bases.cpp:1102-1124)
	const double lam = 57692.30769230769, mu = 38461.53846153846;
	std::vector<double> lam_e(6, lam), mu_e(6, mu);
	pfa_mesh_desc d;
	std::memset(&d, 0, sizeof d);
	d.struct_size = sizeof d;
	d.material = PFA_NEOHOOKEAN;
	d.n_elements = 6;
	d.n_loc = 4;
	d.n_bases = 8;
	d.n_qp = 1;
	d.conn = conn.data();
	d.quad_weights = w;
	d.ref_grads = rg;
	d.vertices = vert.data();
	d.lambda = lam_e.data();
	d.mu = mu_e.data();
	d.material_stride = 1;
	pfa_handle *h = nullptr;
	if (pfa_create(&d, &h) != PFA_OK)
		return std::printf("pfa_create: %s\n", pfa_last_error(nullptr)), 4;
	int64_t ndof, nnz;
	pfa_sizes(h, nullptr, &ndof, &nnz);
	const int32_t *outer, *inner;
	pfa_pattern(h, &nnz, &outer, &inner);
	std::vector<double> x(ndof), g(ndof), v(nnz);
	for (int64_t k = 0; k < ndof; ++k)
		x[k] = 0.03 * std::sin(1.0 + 0.7 * double(k));
	double energy = 0;
	if (pfa_grad_hess(h, x.data(), 0, &energy, g.data(), v.data()) != PFA_OK)
		return std::printf("pfa_grad_hess: %s\n", pfa_last_error(h)), 5;
	// dense copy, symmetry, rigid translations in the null space of H and of the forces
	std::vector<double> H(ndof * ndof, 0.0);
	for (int64_t c = 0; c < ndof; ++c)
		for (int32_t k = outer[c]; k < outer[c + 1]; ++k)
			H[size_t(inner[k]) * ndof + c] = v[k];
	double asym = 0, hmax = 0, fsum = 0, gmax = 0;
	for (int64_t r = 0; r < ndof; ++r)
		for (int64_t c = 0; c < ndof; ++c)
		{
			asym = std::fmax(asym, std::fabs(H[r * ndof + c] - H[c * ndof + r]));
			hmax = std::fmax(hmax, std::fabs(H[r * ndof + c]));
		}
	for (int dd = 0; dd < 3; ++dd)
	{
		double s = 0;
		for (int64_t n = 0; n < 8; ++n)
			s += g[n * 3 + dd];
		fsum = std::fmax(fsum, std::fabs(s));
	}
	for (int64_t k = 0; k < ndof; ++k)
		gmax = std::fmax(gmax, std::fabs(g[k]));
	std::printf("energy %.12e, nnz %lld, asymmetry %.2e of %.2e, force sum %.2e of %.2e\n", energy, (long long)nnz, asym, hmax, fsum, gmax);
	if (!(energy > 0) || asym > 1e-12 * hmax || fsum > 1e-12 * gmax)
		return 6;
	// ---- two ranks, owner-computes: both handles on device 0 ----
	double e_sum = 0, worst = 0;
	int cols = 0;
	for (int r = 0; r < 2; ++r)
	{
		int32_t n_own, n_ghost, n_local, n_owned;
		pfa_partition_sizes(parts[r], &n_own, &n_ghost, &n_local, &n_owned);
		const int32_t *el = pfa_partition_elements(parts[r]), *l2g = pfa_partition_local_to_global(parts[r]);
		const uint8_t *own = pfa_partition_owned(parts[r]);
		const int nt = n_own + n_ghost;
		std::vector<double> vloc(size_t(nt) * 12), lam_l(nt, lam), mu_l(nt, mu), xl(size_t(n_local) * 3);
		for (int t = 0; t < nt; ++t)
			std::memcpy(&vloc[size_t(t) * 12], &vert[size_t(el[t]) * 12], 12 * sizeof(double));
		for (int n = 0; n < n_local; ++n)
			for (int dd = 0; dd < 3; ++dd)
				xl[n * 3 + dd] = x[l2g[n] * 3 + dd];
		pfa_mesh_desc p = d;
		p.n_elements = n_own;
		p.n_ghost_elements = n_ghost;
		p.n_bases = n_local;
		p.conn = pfa_partition_conn(parts[r]);
		p.vertices = vloc.data();
		p.lambda = lam_l.data();
		p.mu = mu_l.data();
		p.owned_nodes = own;
		p.flags = PFA_FLAG_GHOST_GEOMETRY;
		pfa_handle *hr = nullptr;
		if (pfa_create(&p, &hr) != PFA_OK)
			return std::printf("pfa_create (rank %d): %s\n", r, pfa_last_error(nullptr)), 7;
		int64_t ndl, nzl;
		pfa_sizes(hr, nullptr, &ndl, &nzl);
		const int32_t *ol, *il;
		pfa_pattern(hr, &nzl, &ol, &il);
		std::vector<double> gl(ndl, 0.0), vl(nzl, 0.0);
		double er = 0;
		if (pfa_grad_hess(hr, xl.data(), 0, &er, gl.data(), vl.data()) != PFA_OK)
			return std::printf("pfa_grad_hess (rank %d): %s\n", r, pfa_last_error(hr)), 8;
		e_sum += er;
		for (int n = 0; n < n_local; ++n)
		{
			if (!own[n])
				continue;
			for (int dd = 0; dd < 3; ++dd)
			{
				const int64_t cl = int64_t(n) * 3 + dd, cg = int64_t(l2g[n]) * 3 + dd;
				worst = std::fmax(worst, std::fabs(gl[cl] - g[cg]) / gmax);
				for (int32_t k = ol[cl]; k < ol[cl + 1]; ++k)
				{
					const int64_t rg_ = int64_t(l2g[il[k] / 3]) * 3 + il[k] % 3;
					worst = std::fmax(worst, std::fabs(vl[k] - H[rg_ * ndof + cg]) / hmax);
				}
				++cols;
			}
		}
		pfa_destroy(hr);
	}
	std::printf("two ranks: %d owned columns, worst difference %.2e, energy %.12e\n", cols, worst, e_sum);
	pfa_destroy(h);
	for (auto *p : parts)
		pfa_partition_destroy(p);
	return (cols == 24 && worst <= 1e-12 && std::fabs(e_sum - energy) <= 1e-12 * energy) ? 0 : 9;
}
