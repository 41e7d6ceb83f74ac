// Tile planner of the job-list kernel (gemm_mega.cu): which CTA pair ("unit") executes which tile, in which order.
//
// Why: the kernel's default is a static stride over the tiles in job order (unit u takes tiles u, u + U, ...).  That is
// blind to how long a tile takes: at large batches the weight-gradient tiles contract over the whole batch (16 384
// rows = 256 k-blocks, > 100 us each) while the tiles of the dgrad chain take 4-8 us, and every chain tile that the
// stride puts behind a weight-gradient tile on the same unit waits for it -- its row block, and everything downstream
// of it, arrives 100 us late (tools/mega_trace.py at B = 16384: the backward spans 1124 us for 809 us of MMA work per
// unit).  The planner list-schedules the tiles on a model of a unit (in-order MMA issue, the epilogue warps, two TMEM
// accumulator stages) with per-tile costs calibrated on the measured timelines, and hands every unit an explicit tile
// list.  Highest priority = longest remaining dependency chain (HLFET): chain tiles go first, weight-gradient tiles
// (no dependents) fill the units the chain leaves idle, and the longest of them start first.
//
// Deadlock freedom (the units spin on each other's progress): the planner assigns a tile only once every tile it
// depends on has been assigned, and every unit executes its list in assignment order.  So the sequence of assignments
// is a total order in which all dependencies point backwards and every unit's list is increasing: the earliest
// unfinished tile of that order is always runnable -- the argument of the static stride, for any timing whatsoever.
// The cost model only decides how GOOD the plan is, never whether it is valid.
//
// Pure host C++ (no CUDA): tests/csrc/mega_plan_check.cpp drives it on the CPU.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <functional>
#include <queue>
#include <vector>

namespace neraf {
namespace plan {

enum EpilogueKind { EPI_ACT_BF16 = 0, EPI_DGRAD = 1, EPI_WGRAD = 2, EPI_ROWS = 3, EPI_ROWS_TANH = 4 };

struct PlanJob {
  int num_m, num_n;      // tiles: row blocks x column tiles
  int kb;                // k-blocks of 64 per tile
  int bn;                // tile width
  int a_mn, b_mn;
  int wait_job, wait_all;
  int kind;              // EpilogueKind
};

constexpr uint32_t kLocalBits = 20;                     // code = job << 20 | local tile (row-block major)
constexpr uint32_t kLocalMask = (1u << kLocalBits) - 1;

struct Plan {
  std::vector<int32_t> unit_off;     // units + 1 offsets into codes
  std::vector<uint32_t> codes;       // per unit, in execution order
  double makespan = 0.0;             // predicted, us
};

// ---- cost model (us), calibrated on tools/mega_trace.py timelines of the field's forward / backward on a B200:
// a k-block streams from L2 in 0.36 us for tiles <= 128 columns and 0.43 us for 256 columns, MN-major operands
// (dgrad: B; weight gradients: both) are slower; the epilogue costs ~1 us per 64 columns (more with a gate mask and
// column sums, much more for rows TMA cannot tile), publishing a tile ~1.1 us.
inline double mma_us(const PlanJob& j) {
  double kb;
  if (j.bn >= 256) kb = (j.a_mn && j.b_mn) ? 0.49 : (j.b_mn ? 0.465 : 0.43);
  else if (j.bn >= 128) kb = (j.a_mn && j.b_mn) ? 0.43 : (j.b_mn ? 0.375 : 0.36);
  else kb = 0.35;
  return kb * j.kb;
}
inline double epilogue_us(const PlanJob& j) {
  static const double per64[5] = {1.0, 1.3, 1.0, 2.1, 3.4};
  const int k = j.kind < 0 || j.kind > 4 ? 0 : j.kind;
  return per64[k] * (j.bn / 64.0);
}
constexpr double kPublishUs = 1.1;
constexpr double kDependencyUs = 0.8;       // producer's publish -> consumer's first MMA
constexpr double kStageDrain = 0.85;        // fraction of the epilogue after which the TMEM stage is handed back

enum Policy { STATIC_STRIDE = 0, CRITICAL_PATH = 1, ROW_BLOCK = 2 };

namespace detail {
struct Unit {
  double mma_free = 0.0, epi_free = 0.0, stage_free[2] = {0.0, 0.0};
  int count = 0;
};
struct Sim {
  const PlanJob* jobs; int n; int units;
  std::vector<double> tm, te;
  std::vector<Unit> u;
  std::vector<std::vector<uint32_t>> lists;
  std::vector<int> group_off;                 // job -> first (job, row block) group
  std::vector<int> group_left; std::vector<double> group_done;
  std::vector<int> job_groups_left; std::vector<double> job_done;
  Sim(const PlanJob* j, int n_, int units_) : jobs(j), n(n_), units(units_), tm(n_), te(n_), u(units_), lists(units_),
                                               group_off(n_ + 1, 0), job_groups_left(n_), job_done(n_, 0.0) {
    for (int i = 0; i < n; ++i) {
      tm[i] = mma_us(j[i]); te[i] = epilogue_us(j[i]);
      group_off[i + 1] = group_off[i] + j[i].num_m;
      job_groups_left[i] = j[i].num_m;
    }
    group_left.resize(group_off[n]); group_done.assign(group_off[n], 0.0);
    for (int i = 0; i < n; ++i)
      for (int m = 0; m < j[i].num_m; ++m) group_left[group_off[i] + m] = j[i].num_n;
  }
  // run tile (job i, any tile of it) on unit x, operands ready at `ready`; returns the time its results are published
  double run(int x, int i, uint32_t code, double ready) {
    Unit& w = u[x];
    const int st = w.count & 1;
    const double start = std::max(std::max(w.mma_free, w.stage_free[st]), ready);
    const double mma_done = start + tm[i];
    const double e0 = std::max(mma_done, w.epi_free);
    const double fin = e0 + te[i] + kPublishUs;
    w.mma_free = mma_done;
    w.stage_free[st] = e0 + kStageDrain * te[i];
    w.epi_free = fin;
    ++w.count;
    lists[x].push_back(code);
    return fin;
  }
  double decision_time(int x) const { return std::max(u[x].mma_free, u[x].stage_free[u[x].count & 1]); }
  double makespan() const {
    double t = 0.0;
    for (const Unit& w : u) t = std::max(t, w.epi_free);
    return t;
  }
};
}  // namespace detail

// Tiles of job i are numbered row-block major: local = mt * num_n + nt (what the kernel's epilogue expects).
inline Plan make_plan(const PlanJob* jobs, int n, int units, Policy policy) {
  using detail::Sim;
  Sim s(jobs, n, units);
  std::vector<std::vector<int>> dependents(n);
  for (int i = 0; i < n; ++i)
    if (jobs[i].wait_job >= 0) dependents[jobs[i].wait_job].push_back(i);
  // remaining dependency chain behind a tile of job i (HLFET priority)
  std::vector<double> cp(n, 0.0);
  for (int i = n - 1; i >= 0; --i) {
    double down = 0.0;
    for (int d : dependents[i]) down = std::max(down, cp[d] + kDependencyUs);
    cp[i] = s.tm[i] + s.te[i] + kPublishUs + down;
  }
  struct Item { double ready; int job, mt, nt; };
  std::vector<Item> released;
  auto finish_tile = [&](int i, int mt, double fin) {
    const int g = s.group_off[i] + mt;
    s.group_done[g] = std::max(s.group_done[g], fin);
    if (--s.group_left[g] != 0) return;
    s.job_done[i] = std::max(s.job_done[i], s.group_done[g]);
    --s.job_groups_left[i];
    for (int d : dependents[i]) {
      if (!jobs[d].wait_all) {
        if (mt < jobs[d].num_m)
          for (int nt = 0; nt < jobs[d].num_n; ++nt) released.push_back(Item{s.group_done[g] + kDependencyUs, d, mt, nt});
      } else if (s.job_groups_left[i] == 0) {
        for (int m2 = 0; m2 < jobs[d].num_m; ++m2)
          for (int nt = 0; nt < jobs[d].num_n; ++nt) released.push_back(Item{s.job_done[i] + kDependencyUs, d, m2, nt});
      }
    }
  };
  auto code_of = [&](int i, int mt, int nt) { return ((uint32_t)i << kLocalBits) | (uint32_t)(mt * jobs[i].num_n + nt); };

  if (policy == STATIC_STRIDE) {          // the kernel's default order, simulated (for comparison; no table is needed)
    long long pos = 0;
    for (int i = 0; i < n; ++i)
      for (int mt = 0; mt < jobs[i].num_m; ++mt)
        for (int nt = 0; nt < jobs[i].num_n; ++nt, ++pos) {
          const PlanJob& j = jobs[i];
          double r = 0.0;
          if (j.wait_job >= 0)
            r = (j.wait_all ? s.job_done[j.wait_job] : s.group_done[s.group_off[j.wait_job] + mt]) + kDependencyUs;
          const double fin = s.run((int)(pos % units), i, code_of(i, mt, nt), r);
          finish_tile(i, mt, fin);
        }
    released.clear();
  } else {
    // priority key (smaller = first): CRITICAL_PATH: longest remaining chain, then row block; ROW_BLOCK: jobs with
    // dependents first, then row block (a row block runs through the chain as early as possible), then depth
    std::vector<int> depth(n, 0);
    for (int i = 0; i < n; ++i) depth[i] = jobs[i].wait_job >= 0 ? depth[jobs[i].wait_job] + 1 : 0;
    struct Key { double a; int b, c, d, e; };
    auto key_of = [&](int i, int mt, int nt) {
      if (policy == CRITICAL_PATH) return Key{-cp[i], mt, i, nt, 0};
      return Key{dependents[i].empty() ? 1.0 : 0.0, mt, -depth[i], i, nt};
    };
    auto key_less = [](const Key& x, const Key& y) {
      if (x.a != y.a) return x.a < y.a;
      if (x.b != y.b) return x.b < y.b;
      if (x.c != y.c) return x.c < y.c;
      if (x.d != y.d) return x.d < y.d;
      return x.e < y.e;
    };
    struct Ready { Key k; Item it; };
    auto ready_cmp = [&](const Ready& x, const Ready& y) { return key_less(y.k, x.k); };       // min-heap on the key
    auto future_cmp = [](const Item& x, const Item& y) {
      if (x.ready != y.ready) return x.ready > y.ready;
      if (x.job != y.job) return x.job > y.job;
      if (x.mt != y.mt) return x.mt > y.mt;
      return x.nt > y.nt;
    };
    std::priority_queue<Ready, std::vector<Ready>, decltype(ready_cmp)> ready(ready_cmp);
    std::priority_queue<Item, std::vector<Item>, decltype(future_cmp)> future(future_cmp);
    long long total = 0;
    for (int i = 0; i < n; ++i) {
      total += (long long)jobs[i].num_m * jobs[i].num_n;
      if (jobs[i].wait_job < 0)
        for (int mt = 0; mt < jobs[i].num_m; ++mt)
          for (int nt = 0; nt < jobs[i].num_n; ++nt) future.push(Item{0.0, i, mt, nt});
    }
    typedef std::pair<double, int> UT;
    std::priority_queue<UT, std::vector<UT>, std::greater<UT>> free_units;
    for (int x = 0; x < units; ++x) free_units.push(UT(0.0, x));
    long long assigned = 0;
    while (assigned < total) {
      const UT top = free_units.top();
      free_units.pop();
      const double t = top.first;
      const int x = top.second;
      while (!future.empty() && future.top().ready <= t + 1e-9) {
        const Item it = future.top();
        future.pop();
        ready.push(Ready{key_of(it.job, it.mt, it.nt), it});
      }
      if (ready.empty()) {
        if (future.empty()) break;                    // cannot happen for a valid job list (dependencies point backwards)
        free_units.push(UT(future.top().ready, x));   // idle until the next tile becomes runnable
        continue;
      }
      const Item it = ready.top().it;
      ready.pop();
      const double fin = s.run(x, it.job, code_of(it.job, it.mt, it.nt), it.ready);
      ++assigned;
      finish_tile(it.job, it.mt, fin);
      for (const Item& r : released) future.push(r);
      released.clear();
      free_units.push(UT(s.decision_time(x), x));
    }
  }
  Plan p;
  p.makespan = s.makespan();
  p.unit_off.resize(units + 1);
  p.unit_off[0] = 0;
  for (int x = 0; x < units; ++x) p.unit_off[x + 1] = p.unit_off[x] + (int32_t)s.lists[x].size();
  p.codes.reserve(p.unit_off[units]);
  for (int x = 0; x < units; ++x) p.codes.insert(p.codes.end(), s.lists[x].begin(), s.lists[x].end());
  return p;
}

}  // namespace plan
}  // namespace neraf
