/*
 * TEST INFRASTRUCTURE ONLY - CPU restatement of scikit-fmm's distance marcher, the third-party routine behind the
 * reference's geodesic distance field:
 *
 *     dd = skfmm.distance(traversible_ma, dx=1)          nav/agent/agent_state.py:391 (and utils/fmm_planner.py:64, 72)
 *
 * scikit-fmm is pinned by the reference at 2019.1.30 (peanut.Dockerfile:8) and is NOT available in this image (no wheel,
 * no network), so this file restates its published algorithm - skfmm/base_marcher.cpp (initalizeFrozen, initalizeNarrow,
 * solve, cleanUp), skfmm/distance_marcher.cpp (updatePointOrderTwo, solveQuadratic), skfmm/heap.cpp (binary min-heap with
 * back pointers) and the masked-array handling of skfmm/pfmm.py (pre_process_args / post_process_result) - for the one
 * configuration the reference uses: 2-D, dx = 1, order = 2, narrow = 0, not periodic, masked input.
 * PARITY UNPINNED: no reference-held vector exists for this routine; tests check the restatement against closed-form
 * distances (free space: Euclidean to second-order accuracy) and structural invariants, and the CUDA solver against it.
 *
 *   phi   [h*w]  double   level-set function (the reference passes 1 everywhere, 0 at the agent's cell)
 *   mask  [h*w]  uint8    1 = masked (not traversible), 0 = free
 *   out   [h*w]  double   distance; cells that are masked or never reached hold DBL_MAX (pfmm.py masks exactly those)
 * returns 0.  A non-positive discriminant makes solveQuadratic return 0, which solve() treats as "no update" (its
 * `if (d)` tests); only the travel-time marcher raises on it.
 * Least certain points of this restatement (from the published source, not re-checked against the library): the
 * second-order condition uses non-strict comparisons with 0 (`value1 >= 0` / `value1 <= 0`), so the neighbours of a cell
 * with phi == 0 exactly can take the second-order branch across it (values 1/3 instead of 1 on the side whose opposite
 * neighbour froze first); heap ties are resolved by a textbook binary heap with strict comparisons.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

enum { FAR = 0, NARROW = 1, FROZEN = 2, MASK = 3 };

typedef struct {
  int h, w, size;
  const double* phi;
  double* dist;
  signed char* flag;
  /* heap.cpp: binary min-heap over (|distance|, address) with back pointers so that a narrow-band value can be changed */
  int* heap;     /* heap position -> address */
  int* heappos;  /* address -> heap position (-1 when not in the heap) */
  double* key;   /* address -> |distance| */
  int heap_n;
  int error;
} fmm_t;

static int nbr(const fmm_t* m, int cur, int dim, int dir, int flag) {
  /* base_marcher.h _getN: neighbour of `cur` along dim (0 = rows, 1 = columns) at offset dir, or -1 when outside the
   * grid or when the neighbour's flag equals `flag` */
  const int shift = dim == 0 ? m->w : 1;
  const int extent = dim == 0 ? m->h : m->w;
  const int coord = dim == 0 ? cur / m->w : cur % m->w;
  const int nc = coord + dir;
  if (nc >= extent || nc < 0) return -1;
  const int na = cur + dir * shift;
  if (m->flag[na] == flag) return -1;
  return na;
}

static void heap_swap(fmm_t* m, int a, int b) {
  const int ia = m->heap[a], ib = m->heap[b];
  m->heap[a] = ib, m->heap[b] = ia;
  m->heappos[ib] = a, m->heappos[ia] = b;
}
static void sift_up(fmm_t* m, int p) {
  while (p > 0) {
    const int parent = (p - 1) / 2;
    if (m->key[m->heap[p]] < m->key[m->heap[parent]]) heap_swap(m, p, parent), p = parent;
    else break;
  }
}
static void sift_down(fmm_t* m, int p) {
  for (;;) {
    int c = 2 * p + 1;
    if (c >= m->heap_n) break;
    if (c + 1 < m->heap_n && m->key[m->heap[c + 1]] < m->key[m->heap[c]]) c += 1;
    if (m->key[m->heap[c]] < m->key[m->heap[p]]) heap_swap(m, p, c), p = c;
    else break;
  }
}
static void heap_push(fmm_t* m, int addr, double k) {
  m->key[addr] = k;
  m->heap[m->heap_n] = addr;
  m->heappos[addr] = m->heap_n;
  m->heap_n += 1;
  sift_up(m, m->heap_n - 1);
}
static void heap_set(fmm_t* m, int addr, double k) {
  const double old = m->key[addr];
  m->key[addr] = k;
  if (k < old) sift_up(m, m->heappos[addr]);
  else sift_down(m, m->heappos[addr]);
}
static int heap_pop(fmm_t* m, double* k) {
  const int addr = m->heap[0];
  *k = m->key[addr];
  m->heap_n -= 1;
  if (m->heap_n > 0) {
    m->heap[0] = m->heap[m->heap_n];
    m->heappos[m->heap[0]] = 0;
    sift_down(m, 0);
  }
  m->heappos[addr] = -1;
  return addr;
}

/* distance_marcher.cpp solveQuadratic */
static double solve_quadratic(fmm_t* m, int i, double a, double b, double c) {
  c -= 1;
  const double det = b * b - 4 * a * c;
  if (det > 0) {
    if (m->phi[i] > DBL_EPSILON) return (-b + sqrt(det)) / 2.0 / a;
    return (-b - sqrt(det)) / 2.0 / a;
  }
  return 0.0; /* r0 stays 0: the callers in solve() skip the update ("if (d)") */
}

/* distance_marcher.cpp updatePointOrderTwo (dx = 1: idx2 = 1) */
static double update_point(fmm_t* m, int i) {
  const double aa = 9.0 / 4.0, one_third = 1.0 / 3.0;
  double a = 0, b = 0, c = 0;
  for (int dim = 0; dim < 2; ++dim) {
    double value1 = DBL_MAX, value2 = DBL_MAX;
    for (int j = -1; j < 2; j += 2) {
      const int na = nbr(m, i, dim, j, MASK);
      if (na != -1 && m->flag[na] == FROZEN) {
        if (fabs(m->dist[na]) < fabs(value1)) {
          value1 = m->dist[na];
          const int na2 = nbr(m, i, dim, j * 2, MASK);
          if (na2 != -1 && m->flag[na2] == FROZEN &&
              ((m->dist[na2] <= value1 && value1 >= 0) || (m->dist[na2] >= value1 && value1 <= 0)))
            value2 = m->dist[na2];
          else
            value2 = DBL_MAX;
        }
      }
    }
    if (value2 < DBL_MAX) {
      const double tp = one_third * (4 * value1 - value2);
      a += aa;
      b -= 2 * aa * tp;
      c += aa * tp * tp;
    } else if (value1 < DBL_MAX) {
      a += 1;
      b -= 2 * value1;
      c += value1 * value1;
    }
  }
  return solve_quadratic(m, i, a, b, c);
}

int pn_oracle_fmm_distance(const double* phi, const unsigned char* mask, int h, int w, double* out) {
  fmm_t m;
  memset(&m, 0, sizeof(m));
  m.h = h, m.w = w, m.size = h * w, m.phi = phi, m.dist = out;
  m.flag = (signed char*)malloc(m.size);
  m.heap = (int*)malloc(sizeof(int) * m.size);
  m.heappos = (int*)malloc(sizeof(int) * m.size);
  m.key = (double*)malloc(sizeof(double) * m.size);
  for (int i = 0; i < m.size; ++i) {
    m.flag[i] = mask[i] ? MASK : FAR;
    m.dist[i] = DBL_MAX;
    m.heappos[i] = -1;
  }
  /* initalizeFrozen, part 1: points exactly on the zero level set */
  for (int i = 0; i < m.size; ++i)
    if (m.flag[i] != MASK && phi[i] == 0.0) m.flag[i] = FROZEN, m.dist[i] = 0.0;
  /* part 2: far points whose neighbour has the opposite sign (linear interpolation of the crossing) */
  for (int i = 0; i < m.size; ++i) {
    if (m.flag[i] != FAR) continue;
    double ld[2] = {0, 0};
    int borders = 0;
    for (int dim = 0; dim < 2; ++dim)
      for (int j = -1; j < 2; j += 2) {
        const int na = nbr(&m, i, dim, j, MASK);
        if (na != -1 && phi[i] * phi[na] < 0) {
          borders = 1;
          const double d = phi[i] / (phi[i] - phi[na]);
          if (ld[dim] == 0 || ld[dim] > d) ld[dim] = d;
        }
      }
    if (borders) {
      double dsum = 0;
      for (int dim = 0; dim < 2; ++dim)
        if (ld[dim] > 0) dsum += 1 / ld[dim] / ld[dim];
      m.dist[i] = phi[i] < 0 ? -sqrt(1 / dsum) : sqrt(1 / dsum);
      m.flag[i] = FROZEN;
    }
  }
  /* initalizeNarrow */
  for (int i = 0; i < m.size; ++i) {
    if (m.flag[i] != FAR) continue;
    for (int dim = 0; dim < 2; ++dim)
      for (int j = -1; j < 2; j += 2) {
        const int na = nbr(&m, i, dim, j, MASK);
        if (na != -1 && m.flag[na] == FROZEN && m.flag[i] == FAR) {
          m.flag[i] = NARROW;
          const double d = update_point(&m, i);
          m.dist[i] = d;
          heap_push(&m, i, fabs(d));
        }
      }
  }
  /* solve */
  while (m.heap_n > 0 && !m.error) {
    double value;
    const int addr = heap_pop(&m, &value);
    m.flag[addr] = FROZEN;
    for (int dim = 0; dim < 2; ++dim)
      for (int j = -1; j < 2; j += 2) {
        const int na = nbr(&m, addr, dim, j, FROZEN);
        if (na != -1 && m.flag[na] != FROZEN) {
          if (m.flag[na] == NARROW) {
            const double d = update_point(&m, na);
            if (d) {
              heap_set(&m, na, fabs(d));
              m.dist[na] = d;
            }
          } else if (m.flag[na] == FAR) {
            const double d = update_point(&m, na);
            if (d) {
              m.dist[na] = d;
              m.flag[na] = NARROW;
              heap_push(&m, na, fabs(d));
            }
          }
        }
        /* order 2: the point two cells away sees the newly frozen value through its second-order stencil */
        const int local = nbr(&m, addr, dim, j, MASK);
        if (local != -1 && m.flag[local] == FROZEN) {
          const int na2 = nbr(&m, addr, dim, j * 2, FROZEN);
          if (na2 != -1 && m.flag[na2] == NARROW) {
            const double d = update_point(&m, na2);
            if (d) {
              heap_set(&m, na2, fabs(d));
              m.dist[na2] = d;
            }
          }
        }
      }
  }
  /* cleanUp */
  for (int i = 0; i < m.size; ++i)
    if (m.flag[i] != FROZEN) m.dist[i] = DBL_MAX;
  const int err = m.error;
  free(m.flag), free(m.heap), free(m.heappos), free(m.key);
  return err;
}
