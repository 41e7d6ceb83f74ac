"""Reads a NERAF_MEGA_TRACE file (per-tile globaltimer stamps of the job-list kernel) and prints, per launch and job,
where the time goes: dependency wait, operand latency, MMA, epilogue, signalling.

    NERAF_MEGA_TRACE=gpurun_out/trace.bin python tools/bench_jobs.py ; python tools/mega_trace.py gpurun_out/trace.bin
"""
import struct
import sys

import numpy as np

SLOTS = ["dep", "loaded", "mma_start", "mma_first", "mma_done", "epi_start", "epi_stored", "epi_done"]


def read(path):
    data = open(path, "rb").read()
    off, launches = 0, []
    while off < len(data):
        magic, n_jobs, n_tiles, units = struct.unpack_from("4i", data, off); off += 16
        assert magic in (0x4d454741, 0x4d454732)
        jobs = [struct.unpack_from("8i", data, off + 32 * i) for i in range(n_jobs)]; off += 32 * n_jobs
        # which job every stamp slot belongs to: v2 files say it per slot (an explicit tile plan permutes the tiles)
        job_of = np.zeros(n_tiles, dtype=np.int64)
        for ji, j in enumerate(jobs):
            job_of[j[0]:] = ji
        if magic == 0x4d454732:
            codes = np.frombuffer(data, dtype=np.uint32, count=n_tiles, offset=off); off += 4 * n_tiles
            job_of = np.where(codes == 0xffffffff, job_of, codes >> 20).astype(np.int64)
        raw = np.frombuffer(data, dtype=np.uint64, count=n_tiles * 16, offset=off).reshape(n_tiles, 16).astype(np.float64); off += n_tiles * 128
        launches.append((jobs, units, raw[:, :8].copy(), raw[:, 8:].copy(), job_of))
    return launches


def main():
    launches = read(sys.argv[1])
    which = [int(a) for a in sys.argv[2:]] or range(len(launches))
    for li in which:
        jobs, units, t, ck, job_of = launches[li]
        t = np.where(t == 0, np.nan, t)
        t0 = np.nanmin(t)
        t = (t - t0) / 1e3                                                         # us since the first stamp
        print(f"=== launch {li}: {len(jobs)} jobs, {t.shape[0]} tiles on {units} units, span {np.nanmax(t):.1f} us")
        print("job   M     N     K   bn mn wait tiles | mma_start first..last | epi_done first..last | per tile (us): dep->first  first->mma_done  mma_done->epi_start  epi  signal")
        for ji, (ts, M, N, K, bn, a_mn, b_mn, wait) in enumerate(jobs):
            sel = job_of == ji
            te = ts + int(sel.sum())
            x = t[sel]
            c = {s: x[:, i] for i, s in enumerate(SLOTS)}
            med = lambda v: float(np.nanmedian(v)) if np.isfinite(v).any() else float("nan")   # noqa: E731
            print(f"{ji:2d} {M:5d} {N:5d} {K:5d} {bn:4d} {a_mn}{b_mn} {wait:4d} {te - ts:5d} | "
                  f"{np.nanmin(c['mma_start']):7.1f} {np.nanmax(c['mma_start']):7.1f} | {np.nanmin(c['epi_done']):7.1f} {np.nanmax(c['epi_done']):7.1f} | "
                  f"{med(c['mma_first'] - c['dep']):6.2f} {med(c['mma_done'] - c['mma_first']):6.2f} {med(c['epi_start'] - c['mma_done']):6.2f} "
                  f"{med(c['epi_stored'] - c['epi_start']):6.2f} {med(c['epi_done'] - c['epi_stored']):6.2f}   wait-for-dep(mma_first-mma_start) {med(c['mma_first'] - c['mma_start']):6.2f}")
            k = ck[sel]
            k = np.where(k == 0, np.nan, k)
            print("      tracer-warp cycles: ld %.0f  math %.0f  store %.0f  rest-of-loop %.0f  fence %.0f" % tuple(
                med(k[:, i + 1] - k[:, i]) for i in range(5)))
        # unit utilisation: time with an MMA in flight / span
        busy = np.nansum(t[:, 4] - t[:, 3]) / (units * np.nanmax(t))
        print(f"MMA-issue busy fraction over all units: {busy:.2f}")


if __name__ == "__main__":
    main()


def unit_timeline(path, li, unit):
    """Per-tile stamps (us) of one unit in launch li: where its time goes between consecutive tiles."""
    jobs, units, t, ck, job_of = read(path)[li]
    t = np.where(t == 0, np.nan, t)
    t = (t - np.nanmin(t)) / 1e3
    starts = [j[0] for j in jobs]
    print(f"launch {li} unit {unit}:  tile job |  dep  loaded | mma_start mma_first mma_done | epi_start epi_stored epi_done")
    for tile in range(unit, t.shape[0], units):
        j = max(i for i, s in enumerate(starts) if s <= tile)
        r = t[tile]
        print(f"  {tile:5d} {j:3d} | {r[0]:6.1f} {r[1]:6.1f} | {r[2]:6.1f} {r[3]:6.1f} {r[4]:6.1f} | {r[5]:6.1f} {r[6]:6.1f} {r[7]:6.1f}")
