"""Prototype of the job-list kernel's tile planner: a unit-level simulator (MMA pipe + epilogue warps + two TMEM stages)
and list-scheduling policies, calibrated on tools/mega_trace.py timelines.  Used to choose the policy implemented in
csrc/mega_plan.h."""
import heapq
import math
import sys
from dataclasses import dataclass, field


@dataclass
class Job:
    M: int
    N: int
    K: int
    bn: int
    wait: int = -1
    wait_all: int = 0
    a_mn: int = 0
    b_mn: int = 0
    kind: str = "act"        # act (bf16 + activation + mask), dgrad, wgrad, rows (fp32 rows, non-TMA), rows_tanh
    boost: int = 0

    @property
    def num_m(self): return (self.M + 255) // 256
    @property
    def num_n(self): return (self.N + self.bn - 1) // self.bn
    @property
    def kb(self): return (self.K + 63) // 64


def t_mma(j):
    if j.bn == 256:
        tk = 0.49 if (j.a_mn and j.b_mn) else (0.465 if j.b_mn else 0.43)
    elif j.bn == 128:
        tk = 0.43 if (j.a_mn and j.b_mn) else (0.375 if j.b_mn else 0.36)
    else:
        tk = 0.35
    return j.kb * tk


def t_epi(j):
    per64 = {"act": 1.0, "dgrad": 1.3, "wgrad": 1.0, "rows": 2.1, "rows_tanh": 3.4}[j.kind]
    return per64 * (j.bn / 64), 1.1          # (epilogue, publish)


DEP_LAT = 0.8


def simulate(jobs, units, policy):
    """policy: 'static' (index order, stride), or a key function name for list scheduling.  Returns makespan, per-unit lists."""
    n = len(jobs)
    tm = [t_mma(j) for j in jobs]
    te = [t_epi(j) for j in jobs]
    depth = [0] * n
    has_dep = [False] * n
    for i, j in enumerate(jobs):
        if j.wait >= 0:
            depth[i] = depth[j.wait] + 1
            has_dep[j.wait] = True
    # downstream critical path
    cp = [0.0] * n
    for i in range(n - 1, -1, -1):
        down = max([cp[d] + DEP_LAT for d in range(n) if jobs[d].wait == i], default=0.0)
        cp[i] = tm[i] + te[i][0] + te[i][1] + down
    # unit state
    mma_free = [0.0] * units
    epi_free = [0.0] * units
    stage_free = [[0.0, 0.0] for _ in range(units)]
    count = [0] * units
    lists = [[] for _ in range(units)]
    group_left = {}
    group_done = {}
    for i, j in enumerate(jobs):
        for mt in range(j.num_m):
            group_left[(i, mt)] = j.num_n
            group_done[(i, mt)] = 0.0
    job_groups_left = [j.num_m for j in jobs]
    job_done = [0.0] * n
    done_time = {}

    def run(u, i, mt, nt, ready):
        st = count[u] & 1
        start = max(mma_free[u], stage_free[u][st], ready)
        mma_done = start + tm[i]
        e0 = max(mma_done, epi_free[u])
        e1 = e0 + te[i][0]
        fin = e1 + te[i][1]
        mma_free[u] = mma_done
        stage_free[u][st] = e0 + 0.85 * te[i][0]
        epi_free[u] = fin
        count[u] += 1
        lists[u].append((i, mt, nt, start, fin))
        return fin

    released = []            # callbacks: tiles whose ready time became known

    def finish_tile(i, mt, fin):
        g = (i, mt)
        group_left[g] -= 1
        group_done[g] = max(group_done[g], fin)
        if group_left[g] == 0:
            job_groups_left[i] -= 1
            job_done[i] = max(job_done[i], group_done[g])
            for d in range(n):
                if jobs[d].wait == i:
                    if not jobs[d].wait_all:
                        for nt in range(jobs[d].num_n):
                            released.append((group_done[g] + DEP_LAT, d, mt, nt))
                    elif job_groups_left[i] == 0:
                        for m2 in range(jobs[d].num_m):
                            for nt in range(jobs[d].num_n):
                                released.append((job_done[i] + DEP_LAT, d, m2, nt))

    if policy == "static":
        idx = 0
        order = []
        for i, j in enumerate(jobs):
            for mt in range(j.num_m):
                for nt in range(j.num_n):
                    order.append((i, mt, nt))
        ready_of = {}
        for pos, (i, mt, nt) in enumerate(order):
            j = jobs[i]
            if j.wait < 0:
                r = 0.0
            elif j.wait_all:
                r = job_done[j.wait] + DEP_LAT
            else:
                r = group_done[(j.wait, mt)] + DEP_LAT
            u = pos % units
            fin = run(u, i, mt, nt, r)
            finish_tile(i, mt, fin)
        return max(epi_free), lists

    keyf = {
        "cp": lambda i, mt, nt: (-cp[i], -jobs[i].boost, mt, i, nt),
        "deep": lambda i, mt, nt: (-depth[i] - jobs[i].boost, 0 if has_dep[i] else 1, mt, i, nt),
        "chain_deep": lambda i, mt, nt: (0 if has_dep[i] else 1, -depth[i] - jobs[i].boost, mt, i, nt),
        "chain_cp": lambda i, mt, nt: (0 if has_dep[i] else 1, -cp[i], mt, i, nt),
        "rb": lambda i, mt, nt: (0 if has_dep[i] else 1, mt, -depth[i], i, nt),
    }[policy]
    future, ready = [], []
    for i, j in enumerate(jobs):
        if j.wait < 0:
            for mt in range(j.num_m):
                for nt in range(j.num_n):
                    heapq.heappush(future, (0.0, i, mt, nt))
    unit_heap = [(0.0, u) for u in range(units)]
    heapq.heapify(unit_heap)
    total = sum(j.num_m * j.num_n for j in jobs)
    assigned = 0
    while assigned < total:
        t, u = heapq.heappop(unit_heap)
        while future and future[0][0] <= t + 1e-9:
            r, i, mt, nt = heapq.heappop(future)
            heapq.heappush(ready, (keyf(i, mt, nt), r, i, mt, nt))
        if not ready:
            assert future, "deadlock in the planner"
            heapq.heappush(unit_heap, (future[0][0], u))
            continue
        _, r, i, mt, nt = heapq.heappop(ready)
        fin = run(u, i, mt, nt, r)
        assigned += 1
        finish_tile(i, mt, fin)
        for item in released:
            heapq.heappush(future, item)
        released.clear()
        st = count[u] & 1
        heapq.heappush(unit_heap, (max(mma_free[u], stage_free[u][st]), u))
    return max(epi_free), lists


def choose_bn_fwd(M, N, K, min_bn=64, pairs=74):
    kb = math.ceil(K / 64)
    best, best_t = min_bn, 1e30
    bn = min_bn
    while bn <= 256:
        tiles = math.ceil(M / 256) * math.ceil(N / bn)
        mma = kb * (0.41 if bn == 256 else 0.36)
        epi = 0.9 * (bn / 64) + 0.9
        t = (math.ceil(tiles / pairs) - 1.0) * max(mma, epi) + mma + epi
        if t < best_t * (1.02 if mma < epi else 0.97):
            best_t, best = t, bn
        bn *= 2
    return best


def forward_jobs(B, widths=(5096, 2048, 1024, 1024, 512), E=163, CF=513):
    jobs = []
    k = E
    for li, n in enumerate(widths):
        jobs.append(Job(B, n, k, choose_bn_fwd(B, n, k), wait=li - 1, kind="act"))
        k = n
    jobs.append(Job(B, CF, k, choose_bn_fwd(B, CF, k), wait=len(widths) - 1, kind="rows_tanh"))
    return jobs


def backward_jobs(B, widths=(5096, 2048, 1024, 1024, 512), E=163, CF=513, dw2_first=False):
    jobs = []
    W = widths[-1]
    jobs.append(Job(B, W, CF, choose_bn_fwd(B, W, CF, 128), wait=-1, b_mn=1, kind="dgrad"))
    jobs.append(Job(CF, W, B, 256 if W >= 2048 else choose_bn_fwd(CF, W, B, 128), wait=-1, wait_all=0, a_mn=1, b_mn=1, kind="wgrad"))
    producer = 0
    L = len(widths)
    for i in range(L - 1, -1, -1):
        n_i = widths[i]
        k_i = widths[i - 1] if i > 0 else E
        dz = producer

        def dgrad():
            nonlocal producer
            if i > 0:
                jobs.append(Job(B, k_i, n_i, choose_bn_fwd(B, k_i, n_i, 128), wait=dz, b_mn=1, kind="dgrad"))
                producer = len(jobs) - 1

        def wgrad():
            bn = 256 if k_i >= 2048 else choose_bn_fwd(n_i, k_i, B, 128)
            jobs.append(Job(n_i, k_i, B, bn, wait=dz, wait_all=1, a_mn=1, b_mn=1, kind="wgrad" if i > 0 else "rows"))
        if dw2_first and i == 1:
            wgrad(); dgrad()
        else:
            dgrad(); wgrad()
    return jobs


if __name__ == "__main__":
    for B in (2048, 16384):
        for name, jobs in (("forward", forward_jobs(B)), ("backward", backward_jobs(B))):
            print(f"B={B} {name}: " + "  ".join(f"{j.M}x{j.N}x{j.K}/{j.bn}" for j in jobs))
            work = sum(j.num_m * j.num_n * t_mma(j) for j in jobs) / 74
            for pol in ("static", "cp", "deep", "chain_deep", "chain_cp", "rb"):
                ms, lists = simulate(jobs, 74, pol)
                print(f"   {pol:11s} makespan {ms:8.1f} us   (MMA work / unit {work:.1f} us)")


def job_spans(jobs, lists):
    sp = {}
    for u, l in enumerate(lists):
        for (i, mt, nt, s, f) in l:
            a, b, c = sp.get(i, (1e30, 0.0, 0))
            sp[i] = (min(a, s), max(b, f), c + 1)
    return sp
