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

# ---- render_tc3_kernel: K = 16 tile filled in two steps of 8 positions; lane = 16*khalf + column -------------------------
LBO_A3, LBO_B3, K3 = 2048, 512, 16
seen = {}
for lane in range(32):
    pcol, khalf = lane & 15, lane >> 4
    a_col = (pcol & 7) * 16 + (pcol >> 3) * SBO_A + khalf * LBO_A3
    for kfill in (0, 8):
        for z in range(8):
            off = a_col + (kfill >> 2) * LBO_A3 + 2 * z * SBO_A
            r = 16 * z + pcol
            for i in range(4):
                kp = kfill + 4 * khalf + i
                want = (r % 8) * 16 + (r // 8) * SBO_A + (kp // 4) * LBO_A3 + (kp % 4) * 4
                assert off + 4 * i == want, (lane, kfill, z, i)
                assert (r, kp) not in seen
                seen[(r, kp)] = lane
assert len(seen) == 128 * K3 and max(o for o in [0]) == 0
for q in range(4):
    for z in range(8):
        groups = {(((l & 15) & 7) * 16 + ((l & 15) >> 3) * SBO_A + (l >> 4) * LBO_A3 + 2 * z * SBO_A) // 16 % 8 for l in range(8 * q, 8 * q + 8)}
        assert len(groups) == 8
seenb = {}
for lane in range(32):
    kq, nn_low = lane & 3, lane >> 2
    for kfill in (0, 8):
        for h2 in range(2):
            kp = kfill + kq + 4 * h2
            for e8 in range(3):
                nn = nn_low + 8 * e8
                off = (nn & 7) * 16 + (nn >> 3) * SBO_B + (kp >> 2) * LBO_B3 + (kp & 3) * 4
                assert (nn, kp) not in seenb and off < N * K3 * 4
                seenb[(nn, kp)] = off
    banks = set()
for kfill in (0, 8):
    for h2 in range(2):
        for e8 in range(3):
            banks = {(((l >> 2) + 8 * e8) & 7) * 4 + ((kfill + (l & 3) + 4 * h2) & 3) for l in range(32)}
            assert len(banks) == 32        # the 32 scalar S stores of one instruction hit 32 different banks
assert len(seenb) == 24 * K3
print("tc3 operand layout ok: A 128x16 and B 24x16 elements each written once at canonical offsets; stores conflict-free")
