// CPU check of the job-list kernel's tile planner (neraf_b200/csrc/mega_plan.h): every plan must (1) contain every tile
// exactly once, (2) be executable by units that walk their lists strictly in order and spin on dependencies -- for any
// timing --, and (3) for the acoustic field's own job lists at large batches, beat the static stride on the model.
// usage: mega_plan_check            -> runs all checks, prints the makespans, exit code 0 when everything holds
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../neraf_b200/csrc/mega_plan.h"

using namespace neraf::plan;

static int ceil_div(int a, int b) { return (a + b - 1) / b; }

static PlanJob job(int M, int N, int K, int bn, int wait, int wait_all, int a_mn, int b_mn, int kind) {
  PlanJob j;
  j.num_m = ceil_div(M, 256); j.num_n = ceil_div(N, bn); j.kb = ceil_div(K, 64); j.bn = bn;
  j.a_mn = a_mn; j.b_mn = b_mn; j.wait_job = wait; j.wait_all = wait_all; j.kind = kind;
  return j;
}

// the field's forward / backward lists (csrc/field.cu) with the tile widths the library picks at these batches
static std::vector<PlanJob> forward_jobs(int B, bool wide) {
  const int w[5] = {5096, 2048, 1024, 1024, 512};
  std::vector<PlanJob> v;
  int k = 163;
  for (int i = 0; i < 5; ++i) {
    const int bn = wide ? 256 : (i < 2 ? 256 : (i < 4 ? 128 : 64));
    v.push_back(job(B, w[i], k, bn, i - 1, 0, 0, 0, EPI_ACT_BF16));
    k = w[i];
  }
  v.push_back(job(B, 513, k, wide ? 256 : 64, 4, 0, 0, 0, EPI_ROWS_TANH));
  return v;
}
static std::vector<PlanJob> backward_jobs(int B, bool wide) {
  const int w[5] = {5096, 2048, 1024, 1024, 512};
  std::vector<PlanJob> v;
  v.push_back(job(B, 512, 513, wide ? 256 : 128, -1, 0, 0, 1, EPI_DGRAD));
  v.push_back(job(513, 512, B, 128, -1, 0, 1, 1, EPI_WGRAD));
  int producer = 0;
  for (int i = 4; i >= 0; --i) {
    const int n = w[i], k = i > 0 ? w[i - 1] : 163;
    const int dz = producer;
    if (i > 0) {
      v.push_back(job(B, k, n, (wide || k >= 2048) ? 256 : 128, dz, 0, 0, 1, EPI_DGRAD));
      producer = (int)v.size() - 1;
    }
    v.push_back(job(n, k, B, k >= 2048 ? 256 : 128, dz, 1, 1, 1, i > 0 ? EPI_WGRAD : EPI_ROWS));
  }
  return v;
}

// (1) + (2): returns false (and says why) if the plan is not a valid schedule
static bool valid(const std::vector<PlanJob>& jobs, const Plan& p, int units, const char* what) {
  const int n = (int)jobs.size();
  std::vector<int> off(n + 1, 0), goff(n + 1, 0);
  for (int i = 0; i < n; ++i) { off[i + 1] = off[i] + jobs[i].num_m * jobs[i].num_n; goff[i + 1] = goff[i] + jobs[i].num_m; }
  if ((int)p.codes.size() != off[n] || (int)p.unit_off.size() != units + 1 || p.unit_off[units] != off[n]) {
    fprintf(stderr, "%s: wrong number of tiles (%zu of %d)\n", what, p.codes.size(), off[n]);
    return false;
  }
  std::vector<char> seen(off[n], 0);
  for (uint32_t c : p.codes) {
    const int i = (int)(c >> kLocalBits), t = (int)(c & kLocalMask);
    if (i >= n || t >= jobs[i].num_m * jobs[i].num_n || seen[off[i] + t]) { fprintf(stderr, "%s: bad / repeated tile\n", what); return false; }
    seen[off[i] + t] = 1;
  }
  // greedy execution: a unit's head tile completes as soon as its dependency is complete; monotone, so if this gets
  // stuck every execution gets stuck, and if it finishes none can deadlock
  std::vector<int> head(units), group_done(goff[n], 0), job_groups(n, 0);
  for (int x = 0; x < units; ++x) head[x] = p.unit_off[x];
  int finished = 0;
  bool progress = true;
  while (progress) {
    progress = false;
    for (int x = 0; x < units; ++x) {
      while (head[x] < p.unit_off[x + 1]) {
        const uint32_t c = p.codes[head[x]];
        const int i = (int)(c >> kLocalBits), t = (int)(c & kLocalMask), mt = t / jobs[i].num_n;
        const PlanJob& j = jobs[i];
        bool ok = true;
        if (j.wait_job >= 0) {
          const PlanJob& w = jobs[j.wait_job];
          if (j.wait_all) ok = job_groups[j.wait_job] == w.num_m;
          else ok = group_done[goff[j.wait_job] + mt] == w.num_n;
        }
        if (!ok) break;
        if (++group_done[goff[i] + mt] == j.num_n) ++job_groups[i];
        ++head[x]; ++finished; progress = true;
      }
    }
  }
  if (finished != off[n]) { fprintf(stderr, "%s: deadlock after %d of %d tiles\n", what, finished, off[n]); return false; }
  return true;
}

int main() {
  int bad = 0;
  const int units = 74;
  const Policy pol[3] = {STATIC_STRIDE, CRITICAL_PATH, ROW_BLOCK};
  const char* pname[3] = {"static", "critical-path", "row-block"};
  for (int B : {256, 2048, 4096, 16384, 65536}) {
    for (int dir = 0; dir < 2; ++dir) {
      const std::vector<PlanJob> jobs = dir ? backward_jobs(B, B >= 4096) : forward_jobs(B, B >= 4096);
      double ms[3];
      for (int k = 0; k < 3; ++k) {
        const Plan p = make_plan(jobs.data(), (int)jobs.size(), units, pol[k]);
        ms[k] = p.makespan;
        char what[96];
        snprintf(what, sizeof what, "B=%d %s %s", B, dir ? "backward" : "forward", pname[k]);
        if (!valid(jobs, p, units, what)) ++bad;
      }
      printf("B=%6d %-8s static %8.1f  critical-path %8.1f  row-block %8.1f us\n", B, dir ? "backward" : "forward", ms[0], ms[1], ms[2]);
      if (dir == 1 && B >= 4096 && !(ms[1] < 0.95 * ms[0])) { fprintf(stderr, "B=%d backward: the plan does not beat the stride\n", B); ++bad; }
    }
  }
  // random job lists: chains with side jobs, both dependency kinds, odd unit counts
  std::mt19937 rng(7);
  for (int trial = 0; trial < 200; ++trial) {
    const int n = 1 + (int)(rng() % 12), u = 1 + (int)(rng() % 90);
    std::vector<PlanJob> jobs;
    std::vector<int> rows;
    for (int i = 0; i < n; ++i) {
      int wait = (i > 0 && rng() % 4) ? (int)(rng() % i) : -1;
      int wait_all = wait >= 0 ? (int)(rng() % 2) : 0;
      int M = wait >= 0 && !wait_all ? rows[wait] : 1 + (int)(rng() % 3000);
      const int bn = 64 << (rng() % 3);
      jobs.push_back(job(M, 1 + (int)(rng() % 2000), 1 + (int)(rng() % 4000), bn, wait, wait_all, (int)(rng() % 2), (int)(rng() % 2),
                         (int)(rng() % 5)));
      rows.push_back(M);
    }
    for (int k = 0; k < 3; ++k) {
      const Plan p = make_plan(jobs.data(), n, u, pol[k]);
      char what[64];
      snprintf(what, sizeof what, "random %d %s", trial, pname[k]);
      if (!valid(jobs, p, u, what)) ++bad;
    }
  }
  printf("%s\n", bad ? "FAILED" : "all plans valid");
  return bad ? 1 : 0;
}
