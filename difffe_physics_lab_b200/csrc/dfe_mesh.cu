// Mesh handle: host-side symbolic analysis (once per mesh) and upload of the device views.
//
// Replaces the per-call Python work of the reference around FEMesh: free_nodes()
// (diffhe/mesh.py:127-129), the dense K[free][:,free] gathers (diffhe/solver.py:171) and the
// O(n_D * n_free) lifting loop (solver.py:166-169) become index lists built here once.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <limits>
#include <new>

#include "dfe_internal.h"

namespace dfe {

static thread_local std::string g_err;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
}

template <class T>
static int upload(dfe_mesh* m, const std::vector<T>& h, const T** dptr) {
  void* d = nullptr;
  size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
  DFE_CUDA_OK(cudaMalloc(&d, bytes));
  m->allocs.push_back(d);
  if (!h.empty()) DFE_CUDA_OK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  *dptr = static_cast<const T*>(d);
  return DFE_OK;
}

}  // namespace dfe

using namespace dfe;

extern "C" const char* dfe_last_error(void) { return g_err.c_str(); }
extern "C" int dfe_abi_version(void) { return DFE_ABI_VERSION; }
extern "C" int dfe_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" void dfe_mesh_destroy(dfe_mesh* m) {
  if (!m) return;
  if (m->info.device >= 0 && !m->allocs.empty()) {
    int cur = -1;
    cudaGetDevice(&cur);
    cudaSetDevice(m->info.device);
    for (void* p : m->allocs) cudaFree(p);
    if (m->h_fault) cudaFreeHost(m->h_fault);
    if (cur >= 0) cudaSetDevice(cur);
  }
  delete m;
}

extern "C" int dfe_mesh_get_info(const dfe_mesh* m, dfe_mesh_info* info) {
  DFE_REQUIRE(m && info, "dfe_mesh_get_info: null argument");
  *info = m->info;
  return DFE_OK;
}

extern "C" int dfe_mesh_csr_host(const dfe_mesh* m, int which, const int64_t** rowptr,
                                 const int64_t** col, int64_t* n_rows, int64_t* nnz) {
  DFE_REQUIRE(m && rowptr && col && n_rows && nnz, "dfe_mesh_csr_host: null argument");
  DFE_REQUIRE(which == 0 || which == 1, "dfe_mesh_csr_host: which must be 0 (K) or 1 (K_free)");
  const auto& rp = which ? m->h_rowptr_f : m->h_rowptr;
  const auto& c = which ? m->h_col_f : m->h_col;
  *rowptr = rp.data();
  *col = c.data();
  *n_rows = static_cast<int64_t>(rp.size()) - 1;
  *nnz = static_cast<int64_t>(c.size());
  return DFE_OK;
}

extern "C" int dfe_mesh_free_nodes_host(const dfe_mesh* m, const int64_t** free_nodes,
                                        int64_t* n_free) {
  DFE_REQUIRE(m && free_nodes && n_free, "dfe_mesh_free_nodes_host: null argument");
  *free_nodes = m->h_free.data();
  *n_free = static_cast<int64_t>(m->h_free.size());
  return DFE_OK;
}

extern "C" int dfe_mesh_create(int dim, int64_t n_nodes, int64_t n_el, const double* nodes,
                               const int64_t* elems, int64_t n_dir, const int64_t* dir_idx,
                               const double* dir_val, int device, dfe_mesh** out) {
  return dfe_mesh_create_p(dim, dim + 1, n_nodes, n_el, nodes, elems, n_dir, dir_idx, dir_val, device, out);
}

extern "C" int dfe_mesh_create_p(int dim, int nodes_per_element, int64_t n_nodes, int64_t n_el, const double* nodes,
                                 const int64_t* elems, int64_t n_dir, const int64_t* dir_idx,
                                 const double* dir_val, int device, dfe_mesh** out) {
  DFE_REQUIRE(out, "dfe_mesh_create: out is null");
  *out = nullptr;
  if (dim != 1 && dim != 2) {
    // reference: NotImplementedError("Only 1D and 2D supported") (solver.py:66-67)
    set_error("Only 1D and 2D supported (dim=%d)", dim);
    return DFE_ERR_UNSUPPORTED;
  }
  DFE_REQUIRE(n_nodes >= 1 && n_el >= 0 && n_dir >= 0, "dfe_mesh_create: negative size");
  DFE_REQUIRE(nodes && (elems || n_el == 0) && ((dir_idx && dir_val) || n_dir == 0),
              "dfe_mesh_create: null array");
  const int npe = nodes_per_element;
  DFE_REQUIRE(npe == dim + 1 || npe == (dim + 1) * (dim + 2) / 2,
              "dfe_mesh_create: %d nodes per element in %dD (P1 elements have %d, P2 elements %d)", npe, dim, dim + 1,
              (dim + 1) * (dim + 2) / 2);
  const int64_t lim = std::numeric_limits<int32_t>::max() / 4;
  DFE_REQUIRE(n_nodes < lim && n_el * npe * npe < lim, "dfe_mesh_create: mesh too large for int32 indices");
  for (int64_t i = 0; i < n_el * npe; ++i)
    DFE_REQUIRE(elems[i] >= 0 && elems[i] < n_nodes, "dfe_mesh_create: element %lld references node %lld outside [0,%lld)",
                (long long)(i / npe), (long long)elems[i], (long long)n_nodes);

  dfe_mesh* m = new (std::nothrow) dfe_mesh();
  DFE_REQUIRE(m, "dfe_mesh_create: out of host memory");
  struct Guard {
    dfe_mesh* m;
    ~Guard() { if (m) dfe_mesh_destroy(m); }
  } guard{m};

  const int n = static_cast<int>(n_nodes), ne = static_cast<int>(n_el), nd = static_cast<int>(n_dir);
  // ---- Dirichlet order / free map (mesh.py:127-129)
  std::vector<int> dir_order(n, -1);
  for (int t = 0; t < nd; ++t) {
    DFE_REQUIRE(dir_idx[t] >= 0 && dir_idx[t] < n_nodes, "dfe_mesh_create: Dirichlet node %lld outside [0,%lld)",
                (long long)dir_idx[t], (long long)n_nodes);
    DFE_REQUIRE(dir_order[dir_idx[t]] < 0, "dfe_mesh_create: Dirichlet node %lld listed twice", (long long)dir_idx[t]);
    dir_order[dir_idx[t]] = t;
  }
  std::vector<int> free_nodes, free_rank(n, -1);
  free_nodes.reserve(n - nd);
  for (int i = 0; i < n; ++i)
    if (dir_order[i] < 0) {
      free_rank[i] = static_cast<int>(free_nodes.size());
      free_nodes.push_back(i);
    }
  const int nfree = static_cast<int>(free_nodes.size());

  // ---- node -> element adjacency, ascending element id (= the reference's accumulation order)
  std::vector<int> h_elems(static_cast<size_t>(ne) * npe);
  for (size_t i = 0; i < h_elems.size(); ++i) h_elems[i] = static_cast<int>(elems[i]);
  std::vector<int> adj_ptr(n + 1, 0);
  for (size_t i = 0; i < h_elems.size(); ++i) adj_ptr[h_elems[i] + 1]++;
  for (int i = 0; i < n; ++i) adj_ptr[i + 1] += adj_ptr[i];
  const int nadj = adj_ptr[n];
  std::vector<int> adj_elem(nadj), adj_loc(nadj), fill(adj_ptr.begin(), adj_ptr.end() - 1);
  for (int e = 0; e < ne; ++e)
    for (int q = 0; q < npe; ++q) {
      int p = h_elems[static_cast<size_t>(e) * npe + q];
      adj_elem[fill[p]] = e;
      adj_loc[fill[p]] = q;
      fill[p]++;
    }

  // ---- structural CSR of K: row p = sorted unique nodes of the elements around p
  std::vector<int> rowptr(n + 1, 0), col;
  col.reserve(static_cast<size_t>(nadj) * 2 + n);
  {
    std::vector<int> tmp;
    for (int p = 0; p < n; ++p) {
      tmp.clear();
      for (int a = adj_ptr[p]; a < adj_ptr[p + 1]; ++a)
        for (int q = 0; q < npe; ++q) tmp.push_back(h_elems[static_cast<size_t>(adj_elem[a]) * npe + q]);
      std::sort(tmp.begin(), tmp.end());
      tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
      col.insert(col.end(), tmp.begin(), tmp.end());
      rowptr[p + 1] = static_cast<int>(col.size());
    }
  }
  const int nnz_full = static_cast<int>(col.size());
  std::vector<int> adj_slot(static_cast<size_t>(nadj) * npe);
  for (int p = 0; p < n; ++p)
    for (int a = adj_ptr[p]; a < adj_ptr[p + 1]; ++a)
      for (int q = 0; q < npe; ++q) {
        int c = h_elems[static_cast<size_t>(adj_elem[a]) * npe + q];
        const int* b = col.data() + rowptr[p];
        const int* e = col.data() + rowptr[p + 1];
        adj_slot[static_cast<size_t>(a) * npe + q] = static_cast<int>(std::lower_bound(b, e, c) - col.data());
      }

  // ---- K_free pattern, diagonal, lifting lists (dict order per row)
  std::vector<int> rowptr_f(nfree + 1, 0), col_f, src_f, diag_src(nfree, -1), lift_ptr(nfree + 1, 0), lift_src;
  std::vector<double> lift_g;
  int max_row = 0;
  {
    std::vector<std::pair<int, int>> lift;  // (dict order, full index)
    for (int r = 0; r < nfree; ++r) {
      int p = free_nodes[r];
      lift.clear();
      for (int k = rowptr[p]; k < rowptr[p + 1]; ++k) {
        int c = col[k];
        if (free_rank[c] >= 0) {
          if (c == p) diag_src[r] = k;
          col_f.push_back(free_rank[c]);
          src_f.push_back(k);
        } else {
          lift.emplace_back(dir_order[c], k);
        }
      }
      std::sort(lift.begin(), lift.end());
      for (auto& lk : lift) {
        lift_src.push_back(lk.second);
        lift_g.push_back(dir_val[lk.first]);
      }
      rowptr_f[r + 1] = static_cast<int>(col_f.size());
      lift_ptr[r + 1] = static_cast<int>(lift_src.size());
      max_row = std::max(max_row, rowptr_f[r + 1] - rowptr_f[r]);
    }
  }
  const int nnz_free = static_cast<int>(col_f.size());

  // ---- SELL-32 layout of K_free
  const int n_slices = (nfree + 31) / 32;
  std::vector<int> slice_ptr(n_slices + 1, 0);
  for (int s = 0; s < n_slices; ++s) {
    int w = 0;
    for (int r = 32 * s; r < std::min(nfree, 32 * s + 32); ++r) w = std::max(w, rowptr_f[r + 1] - rowptr_f[r]);
    slice_ptr[s + 1] = slice_ptr[s] + 32 * w;
  }
  const int sell_nnz = slice_ptr[n_slices];
  std::vector<int> sell_col(sell_nnz), sell_src(sell_nnz, -1);
  for (int s = 0; s < n_slices; ++s) {
    int w = (slice_ptr[s + 1] - slice_ptr[s]) / 32;
    for (int l = 0; l < 32; ++l) {
      int r = 32 * s + l;
      for (int k = 0; k < w; ++k) {
        int pos = slice_ptr[s] + 32 * k + l;
        if (r < nfree && k < rowptr_f[r + 1] - rowptr_f[r]) {
          sell_col[pos] = col_f[rowptr_f[r] + k];
          sell_src[pos] = src_f[rowptr_f[r] + k];
        } else {
          sell_col[pos] = std::min(r, std::max(nfree - 1, 0));  // padding: valid index, value 0
        }
      }
    }
  }

  // ---- 1-D chain detection (fused path): elements (e,e+1), h>0, Dirichlet ⊆ {0, n-1}, >=1 of them
  bool chain = (dim == 1 && npe == 2 && ne >= 1 && n == ne + 1 && nd >= 1 && nd <= 2 && nfree >= 1);
  if (chain)
    for (int e = 0; e < ne && chain; ++e)
      chain = h_elems[2 * e] == e && h_elems[2 * e + 1] == e + 1 && nodes[e + 1] > nodes[e] &&
              std::isfinite(nodes[e + 1] - nodes[e]);
  if (chain)
    for (int t = 0; t < nd; ++t) chain = chain && (dir_idx[t] == 0 || dir_idx[t] == n - 1);
  m->chain = chain;
  if (chain) {
    m->bc_left = dir_order[0] >= 0;
    m->bc_right = dir_order[n - 1] >= 0;
    if (m->bc_left) m->g_left = dir_val[dir_order[0]];
    if (m->bc_right) m->g_right = dir_val[dir_order[n - 1]];
    m->lift_left_first = !(m->bc_left && m->bc_right) || dir_order[0] < dir_order[n - 1];
  }

  // ---- rectangle() topology (mesh.py:79-121): node id = row*(gx+1)+col, quad (r,c) -> [a,b,d], [b,c,d] with
  // a = r(gx+1)+c, b = a+1, c = a+gx+2, d = a+gx+1, and exactly the boundary nodes Dirichlet
  if (dim == 2 && npe == 3 && ne >= 2 && h_elems[0] == 0 && h_elems[1] == 1 && h_elems[2] >= 2) {
    const long long gx = h_elems[2] - 1;
    const long long gy = (n % (gx + 1) == 0) ? n / (gx + 1) - 1 : 0;
    bool ok = gx >= 2 && gy >= 2 && static_cast<long long>(ne) == 2 * gx * gy;
    for (long long q = 0; ok && q < gx * gy; ++q) {
      const int a = static_cast<int>((q / gx) * (gx + 1) + q % gx), b = a + 1, c = a + static_cast<int>(gx) + 2, d = c - 1;
      const int* t = h_elems.data() + 6 * q;
      ok = t[0] == a && t[1] == b && t[2] == d && t[3] == b && t[4] == c && t[5] == d;
    }
    if (ok) {
      m->topo_nx = static_cast<int>(gx);
      m->topo_ny = static_cast<int>(gy);
      // range of the geometry (see dfe_mesh::topo_geo_mid); margins are wide enough for any host rounding mode / contraction
      auto mid = [](double v, int lo, int hi) {
        if (v == 0.0) return true;
        int e;
        std::frexp(v, &e);
        return std::isfinite(v) && e > lo && e < hi;
      };
      bool geo = true;
      for (long long q = 0; geo && q < gx * gy; ++q) {
        const long long a = (q / gx) * (gx + 1) + q % gx;
        const long long tri[2][3] = {{a, a + 1, a + gx + 1}, {a + 1, a + gx + 2, a + gx + 1}};
        for (int t = 0; t < 2 && geo; ++t) {
          const double* P0 = nodes + 2 * tri[t][0];
          const double* P1 = nodes + 2 * tri[t][1];
          const double* P2 = nodes + 2 * tri[t][2];
          const double d[6] = {P1[0] - P0[0], P2[0] - P1[0], P0[0] - P2[0], P1[1] - P0[1], P2[1] - P1[1], P0[1] - P2[1]};
          for (int i = 0; i < 6; ++i) geo = geo && mid(d[i], -120, 120);
          const double area = 0.5 * std::fabs(d[0] * (-d[5]) - (-d[2]) * d[3]);
          geo = geo && area != 0.0 && mid(area, -250, 250);
        }
      }
      m->topo_geo_mid = geo;
    }
    ok = ok && nd == 2 * (gx + gy);
    for (long long p = 0; ok && p < n; ++p) {
      const long long r = p / (gx + 1), cc = p % (gx + 1);
      const bool boundary = r == 0 || r == gy || cc == 0 || cc == gx;
      ok = boundary == (dir_order[p] >= 0);
    }
    if (ok) {
      m->grid_nx = static_cast<int>(gx);
      m->grid_ny = static_cast<int>(gy);
    }
  }

  // ---- host copies exposed through the ABI
  m->h_rowptr.assign(rowptr.begin(), rowptr.end());
  m->h_col.assign(col.begin(), col.end());
  m->h_rowptr_f.assign(rowptr_f.begin(), rowptr_f.end());
  m->h_col_f.assign(col_f.begin(), col_f.end());
  m->h_free.assign(free_nodes.begin(), free_nodes.end());

  dfe_mesh_info& I = m->info;
  I.dim = dim;
  I.device = -1;
  I.n_nodes = n;
  I.n_elements = ne;
  I.n_dirichlet = nd;
  I.n_free = nfree;
  I.nnz_full = nnz_full;
  I.nnz_free = nnz_free;
  I.sell_nnz = sell_nnz;
  I.max_row_nnz = max_row;
  I.chain1d = chain ? 1 : 0;

  if (device >= 0) {
    int ndev = dfe_device_count();
    if (device >= ndev) {
      set_error("dfe_mesh_create: CUDA device %d requested but %d visible (this library has no CPU path)", device, ndev);
      return DFE_ERR_CUDA;
    }
    int cur = -1;
    DFE_CUDA_OK(cudaGetDevice(&cur));
    DFE_CUDA_OK(cudaSetDevice(device));
    I.device = device;
    MeshDev& D = m->dev;
    D.dim = dim; D.npe = npe; D.n_nodes = n; D.n_el = ne; D.n_dir = nd; D.n_free = nfree;
    D.nnz_full = nnz_full; D.nnz_free = nnz_free; D.sell_nnz = sell_nnz; D.n_slices = n_slices;
    std::vector<double> h_nodes(nodes, nodes + static_cast<size_t>(n) * dim);
    std::vector<int> h_dir(nd);
    for (int t = 0; t < nd; ++t) h_dir[t] = static_cast<int>(dir_idx[t]);
    std::vector<double> h_dval(dir_val, dir_val + nd);
    int rc = DFE_OK;
#define UP(vec, field) if (rc == DFE_OK) rc = upload(m, vec, &D.field)
    UP(h_nodes, nodes); UP(h_elems, elems); UP(adj_ptr, adj_ptr); UP(adj_elem, adj_elem);
    UP(adj_loc, adj_loc); UP(adj_slot, adj_slot); UP(rowptr, rowptr); UP(col, col);
    UP(free_nodes, free_nodes); UP(free_rank, free_rank); UP(h_dir, dir_idx); UP(h_dval, dir_val);
    UP(rowptr_f, rowptr_f); UP(col_f, col_f); UP(src_f, src_f); UP(diag_src, diag_src);
    UP(lift_ptr, lift_ptr); UP(lift_src, lift_src); UP(lift_g, lift_g); UP(slice_ptr, slice_ptr);
    UP(sell_col, sell_col); UP(sell_src, sell_src);
#undef UP
    m->n_lift = static_cast<int>(lift_src.size());
    m->n_adj = nadj;
    for (int q = 0; q < n; ++q) m->max_adj = std::max(m->max_adj, adj_ptr[q + 1] - adj_ptr[q]);
    if (rc == DFE_OK && chain) {
      // half element lengths and their reciprocals for the fused 1-D kernels (IEEE double on the host)
      std::vector<double> hsv(ne), rhv(ne);
      for (int e = 0; e < ne; ++e) {
        const volatile double h = nodes[e + 1] - nodes[e];   // solver.py:84-85
        hsv[e] = 0.5 * h;
        rhv[e] = 1.0 / hsv[e];
      }
      rc = upload(m, hsv, &m->d_hs);
      if (rc == DFE_OK) rc = upload(m, rhv, &m->d_rh);
      // X_i = sum_{e<i} h_e/2 (coordinate measured in half element lengths), accumulated in long double so that
      // the table is correct to the last bit or two of a double: the pipelined 1-D kernel writes the prefix
      // sums in moment form (W_i = m + s X_i) and needs one consistent, accurate coordinate per node.
      std::vector<double> xv(n);
      long double acc = 0.0L;
      for (int i = 0; i < n; ++i) {
        xv[i] = static_cast<double>(acc);
        if (i < ne) acc += static_cast<long double>(hsv[i]);
      }
      if (rc == DFE_OK) rc = upload(m, xv, &m->d_X);
      if (rc == DFE_OK) {
        void* hf = nullptr;
        void* df = nullptr;
        if (cudaHostAlloc(&hf, 64, cudaHostAllocMapped) == cudaSuccess && cudaHostGetDevicePointer(&df, hf, 0) == cudaSuccess) {
          m->h_fault = static_cast<int*>(hf);
          m->d_fault = static_cast<int*>(df);
          *m->h_fault = 0;
          void* dd = nullptr;
          if (cudaMalloc(&dd, 64) == cudaSuccess && cudaMemset(dd, 0, 64) == cudaSuccess) {
            m->allocs.push_back(dd);
            m->d_fault_dev = static_cast<int*>(dd);
          } else {
            cudaGetLastError();
            set_error("dfe_mesh_create: cannot allocate the device fault word");
            rc = DFE_ERR_CUDA;
          }
        } else {
          cudaGetLastError();
          if (hf) cudaFreeHost(hf);
          set_error("dfe_mesh_create: cannot allocate the mapped fault word");
          rc = DFE_ERR_CUDA;
        }
      }
      m->x_total = xv[n - 1];
    }
    cudaDeviceProp prop;
    if (rc == DFE_OK && cudaGetDeviceProperties(&prop, device) == cudaSuccess) m->sm_count = prop.multiProcessorCount;
    cudaSetDevice(cur);
    if (rc != DFE_OK) return rc;
  }
  guard.m = nullptr;
  *out = m;
  return DFE_OK;
}
