"""CPU emulation of block_sort_radix (csrc/isect.cu, experimental per-tile radix sort): the same chunking, per-warp
histograms, (digit major, warp minor) scan, match-any ranking, pass skipping and equal-depth fix-up, executed lane by
lane in Python.  Checks that the result is the ascending (depth bits, flatten id) order for random segments incl.
duplicate depths.  Development aid for a kernel that has not run on a GPU yet."""
import random

WARPS, LANES = 8, 32


def match_any(digits):
    return [sum(1 << l2 for l2 in range(LANES) if digits[l2] == digits[l]) for l in range(LANES)]


def ffs(x):
    return (x & -x).bit_length()


def block_sort_radix(a, n):
    a = list(a)
    b = [0] * len(a)
    chunk = (((n + WARPS - 1) // WARPS) + 31) & ~31
    src, dst = a, b
    for p in range(4):
        shift = 32 + 8 * p
        hist = [[0] * 256 for _ in range(WARPS)]
        for w in range(WARPS):
            wb = min(n, w * chunk)
            we = min(n, wb + chunk)
            for g0 in range(wb, we, 32):
                d = [(src[g0 + l] >> shift) & 255 if g0 + l < we else 256 + l for l in range(LANES)]
                peers = match_any(d)
                for l in range(LANES):
                    if g0 + l < we and l == ffs(peers[l]) - 1:
                        hist[w][d[l]] += bin(peers[l]).count("1")
        totals = [sum(hist[w][d] for w in range(WARPS)) for d in range(256)]
        skip = any(t == n for t in totals)
        run = 0
        for d in range(256):
            for w in range(WARPS):
                c = hist[w][d]
                hist[w][d] = run
                run += c
        if skip:
            continue
        for w in range(WARPS):
            wb = min(n, w * chunk)
            we = min(n, wb + chunk)
            for g0 in range(wb, we, 32):
                valid = [g0 + l < we for l in range(LANES)]
                keys = [src[g0 + l] if valid[l] else 0 for l in range(LANES)]
                d = [(keys[l] >> shift) & 255 if valid[l] else 256 + l for l in range(LANES)]
                peers = match_any(d)
                pos = [hist[w][d[l]] + bin(peers[l] & ((1 << l) - 1)).count("1") if valid[l] else 0 for l in range(LANES)]
                for l in range(LANES):
                    if valid[l] and l == ffs(peers[l]) - 1:
                        hist[w][d[l]] += bin(peers[l]).count("1")
                for l in range(LANES):
                    if valid[l]:
                        dst[pos[l]] = keys[l]
        src, dst = dst, src
    out = list(src[:n])
    i = 0
    while i < n:  # equal-depth runs: by flatten id
        j = i + 1
        while j < n and out[j] >> 32 == out[i] >> 32:
            j += 1
        out[i:j] = sorted(out[i:j])
        i = j
    return out


def main():
    rng = random.Random(0)
    for trial in range(300):
        n = rng.choice([1, 2, 31, 32, 33, 255, 256, 257, 500, 1000, 2047, 2048, rng.randint(1, 2048)])
        base = rng.choice([0x3F800000, 0x40000000, 0x41200000])  # depths around 1, 2, 10: shared high bytes
        spread = rng.choice([1, 256, 1 << 16, 1 << 24])
        ids = rng.sample(range(1 << 20), n)
        depths = [base + rng.randrange(spread) for _ in range(n)]
        if n > 4:  # exact ties
            for k in range(0, n - 1, 7):
                depths[k + 1] = depths[k]
        keys = [(d << 32) | i for d, i in zip(depths, ids)]
        got = block_sort_radix(keys + [0xFFFFFFFFFFFFFFFF] * 8, n)
        assert got == sorted(keys), (trial, n)
    print("radix emulation: 300 random segments sorted correctly")


if __name__ == "__main__":
    main()
