"""CPU only: shared-memory bank model of the radix-16 row kernels (csrc/kernels_rows_r16.cuh).

The kernels' case for speed is "four to six conflict-free sweeps over the tile"; this script restates
the index expressions of every shared-memory access of the three kernels (forward direction; the
backward kernels touch the same addresses with loads and stores exchanged) and counts, per warp-wide
128-bit access, how many passes the memory pipe needs: a 32-lane x 16-byte request is served a
quarter warp (8 lanes) at a time, and a quarter warp is conflict-free iff its 8 addresses fall into
8 different 16-byte bank groups (address / 16 mod 8) -- or coincide.  Prints the worst and mean
number of wavefronts per quarter-warp request for each access; 1.00 everywhere = conflict-free.

    python tools/r16_bank_model.py
"""
import numpy as np


def wavefronts(addr):
    """addr: [threads] element (16-byte) indices of one instruction, thread-major.  Returns the
    number of wavefronts of each quarter-warp request (max multiplicity of a bank group among
    DISTINCT addresses)."""
    a = np.asarray(addr).reshape(-1, 8)
    out = []
    for q in a:
        uniq = np.unique(q)
        out.append(np.bincount(uniq % 8, minlength=8).max())
    return np.array(out)


def report(name, per_instruction):
    w = np.concatenate([wavefronts(a) for a in per_instruction])
    print("%-58s worst %d  mean %.2f  (%d quarter-warp requests)" % (name, w.max(), w.mean(), len(w)))
    return w.max()


def r16_out(q):
    return 4 * (q & 3) + (q >> 2)


def model_r16():            # NR = 2048, RB = 2, T = 256
    NR, RB, T, M1, S = 2048, 2, 256, 128, 256
    tid = np.arange(T)
    r, m = tid // M1, tid % M1
    worst = 0
    worst = max(worst, report("r16  pass 1 store  row[q0*128 + m]", [r * NR + q0 * M1 + m for q0 in range(16)]))
    q0, t1 = m >> 3, m & 7
    worst = max(worst, report("r16  pass 2 load   blk[t1 + 8 j]", [r * NR + q0 * M1 + t1 + 8 * j for j in range(16)]))
    worst = max(worst, report("r16  pass 2 store  blk[q1*8 + (t1 ^ s)]",
                              [r * NR + q0 * M1 + (q1 << 3) + (t1 ^ ((q1 & 3) | ((r & 1) << 2))) for q1 in range(16)]))
    rr, p = tid % RB, tid // RB
    klow = ((p >> 2) & 15) | ((p & 3) << 4) | ((p >> 6) << 6)
    klow2 = np.where(p == 0, S // 2, S - klow)

    def slot(k, e):
        q1 = k >> 4
        return rr * NR + ((k & 15) << 7) + (q1 << 3) + (e ^ ((q1 & 3) | ((rr & 1) << 2)))
    worst = max(worst, report("r16  pass 3 load   group klow", [slot(klow, e) for e in range(8)]))
    worst = max(worst, report("r16  pass 3 load   group 256 - klow", [slot(klow2, e) for e in range(8)]))
    return worst


def model_r16h():           # NR = 4096, RB = 2, T = 512
    NR, RB, T, M1 = 4096, 2, 512, 256
    tid = np.arange(T)
    r, m = tid // M1, tid % M1
    worst = 0
    worst = max(worst, report("r16h pass 1 store  row[q0*256 + m]", [r * NR + q0 * M1 + m for q0 in range(16)]))
    q0, t1 = m >> 4, m & 15
    worst = max(worst, report("r16h pass 2 load   blk[t1 + 16 j]", [r * NR + q0 * M1 + t1 + 16 * j for j in range(16)]))
    worst = max(worst, report("r16h pass 2 store  blk[q1*16 + (t1 ^ s)]",
                              [r * NR + q0 * M1 + (q1 << 4) + (t1 ^ ((q1 & 3) | ((r & 1) << 2))) for q1 in range(16)]))
    rr, p = tid % RB, tid // RB
    e_, pp = p >> 7, p & 127
    klow = ((pp >> 2) & 15) | ((pp & 3) << 4) | ((pp >> 6) << 6)
    gA = np.where(pp == 0, np.where(e_ == 1, 128, 0), klow)
    gB = np.where(pp == 0, gA, 256 - klow)

    def slot(g, t):
        q1 = g >> 4
        return rr * NR + ((g & 15) << 8) + (q1 << 4) + (t ^ ((q1 & 3) | ((rr & 1) << 2)))
    worst = max(worst, report("r16h pass 3 load   group A, t and t + 8", [slot(gA, t) for t in range(16)]))
    worst = max(worst, report("r16h pass 3 load   group B, t and t + 8", [slot(gB, t) for t in range(16)]))
    return worst


def model_r16w():           # NR = 8192, one row, T = 512
    NR, T = 8192, 512
    t = np.arange(T)
    worst = 0
    worst = max(worst, report("r16w pass 1 store  sm[q0*512 + t]", [q0 * T + t for q0 in range(16)]))
    for h in range(2):
        i = t + T * h
        worst = max(worst, report("r16w pass 2 load/store blk[64 j], item %d" % h,
                                  [(i >> 6) * 512 + (i & 63) + 64 * j for j in range(8)]))
    for h in range(2):
        i = t + T * h
        t2, q1 = i & 7, (i >> 3) & 7
        worst = max(worst, report("r16w pass 3 load   blk[t2 + 8 j], item %d" % h,
                                  [(i >> 3) * 64 + t2 + 8 * j for j in range(8)]))
        worst = max(worst, report("r16w pass 3 store  blk[q2*8 + (t2 ^ s)], item %d" % h,
                                  [(i >> 3) * 64 + (q2 << 3) + (t2 ^ ((q2 & 3) | ((q1 & 1) << 2))) for q2 in range(8)]))
    p = t
    q1 = ((p >> 7) << 1) | ((p >> 2) & 1)
    klow = ((p >> 3) & 15) | (q1 << 4) | ((p & 3) << 7)
    klow2 = np.where(p == 0, 512, 1024 - klow)

    def slot(g, e):
        q1_, q2_ = (g >> 4) & 7, g >> 7
        return ((g & 15) << 9) + (q1_ << 6) + (q2_ << 3) + (e ^ ((q2_ & 3) | ((q1_ & 1) << 2)))
    worst = max(worst, report("r16w pass 4 load   group klow", [slot(klow, e) for e in range(8)]))
    worst = max(worst, report("r16w pass 4 load   group 1024 - klow", [slot(klow2, e) for e in range(8)]))
    # the unit -> group maps must be bijections
    assert sorted(klow.tolist()) == list(range(512))
    return worst


if __name__ == "__main__":
    w = max(model_r16(), model_r16h(), model_r16w())
    print("worst case over all accesses: %d wavefront(s) per quarter-warp request" % w)
