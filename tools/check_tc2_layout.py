"""Host-side check of render_tc2_kernel's operand addressing (splat_forward_tc.cu): every (row, k) element of the A
tile and every (class, k) element of the B tile is written exactly once at its canonical K-major SWIZZLE_NONE offset
   (m % 8) * 16 + (m / 8) * SBO + (k / 4) * LBO + (k % 4) * 4
and the 16-byte A stores of a quarter warp fall into eight different 16-byte bank groups."""
SBO_A, LBO_A, SBO_B, LBO_B, K, N = 128, 2048, 128, 512, 32, 32
seen = {}
for tid in range(128):
    lane, warp = tid & 31, tid >> 5
    pcol = (lane & 7) + 8 * ((lane >> 3) & 1)
    kgrp = (warp & 3) * 2 + (lane >> 4)
    a_col = (pcol & 7) * 16 + (pcol >> 3) * SBO_A + kgrp * LBO_A
    for z in range(8):
        off = a_col + 2 * z * SBO_A
        r = 16 * z + pcol
        for i in range(4):
            k = 4 * kgrp + i
            want = (r % 8) * 16 + (r // 8) * SBO_A + (k // 4) * LBO_A + (k % 4) * 4
            assert off + 4 * i == want, (tid, z, i)
            assert (r, k) not in seen
            seen[(r, k)] = tid
assert len(seen) == 128 * K
for warp in range(4):
    for q in range(4):                      # quarter warps issue together
        for z in range(8):
            groups = set()
            for lane in range(8 * q, 8 * q + 8):
                pcol = (lane & 7) + 8 * ((lane >> 3) & 1)
                kgrp = warp * 2 + (lane >> 4)
                off = (pcol & 7) * 16 + (pcol >> 3) * SBO_A + kgrp * LBO_A + 2 * z * SBO_A
                groups.add((off // 16) % 8)
            assert len(groups) == 8, (warp, q, z, groups)
seenb = {}
for tid in range(128):
    lane, warp = tid & 31, tid >> 5
    for h2 in range(2):
        kk = (lane & 3) + 4 * (warp & 3) + 16 * h2
        for e8 in range(3):
            nn = (lane >> 2) + 8 * e8
            off = (nn & 7) * 16 + (nn >> 3) * SBO_B + (kk >> 2) * LBO_B + (kk & 3) * 4
            assert (nn, kk) not in seenb
            seenb[(nn, kk)] = off
            assert off < N * K * 4
assert len(seenb) == 24 * K
# epilogue rows: thread tid reads TMEM lane tid = row 16*z + column
for tid in range(128):
    assert ((tid >> 4) & 7) * 16 + (tid & 15) == tid
print("tc2 operand layout ok: A 128x32 and B 24x32 elements each written once at canonical offsets; A stores conflict-free")
