"""Index arithmetic of the K-major tensor-core Gram variant (irspack_b200/csrc/wgram_k.cu),
emulated on the CPU: which shared-memory word every producer lane writes, read back with the
K-major SWIZZLE_128B convention of csrc/tc.cuh (`sw128_offset`, the layout score_tc.cu's
GPU-verified producers use).  Guards the arithmetic only; the kernel itself is checked on the
GPU by tests/test_zz_experimental.py."""
import numpy as np

KT, KP = 32, 128  # neighbours per stage, padded feature dimension (wgram_k.cu)


def sw128_offset(r: int, c16: int) -> int:  # csrc/tc.cuh
    return (r >> 3) * 1024 + (r & 7) * 128 + ((c16 ^ (r & 7)) << 4)


def producer_stores(Y: np.ndarray):
    """Replays the store loop of the 4 producer warps of one group for one stage:
    warp pw owns neighbours 8 pw .. 8 pw + 7, lane l the features l, l + 32, l + 64, l + 96."""
    tile = np.full(KP * 128 // 4, np.nan, np.float32)
    writes = np.zeros(tile.shape, dtype=int)
    for pw in range(4):
        for lane in range(32):
            for j in range(4):
                f = lane + 32 * j
                for half in range(2):
                    off = sw128_offset(f, 2 * pw + half)
                    for e in range(4):
                        t = 8 * pw + 4 * half + e
                        tile[off // 4 + e] = Y[t, f]
                        writes[off // 4 + e] += 1
    return tile, writes


def test_kmajor_tile_is_feature_rows_by_neighbour_columns():
    Y = np.random.default_rng(0).standard_normal((KT, KP)).astype(np.float32)
    tile, writes = producer_stores(Y)
    assert (writes == 1).all()  # the 16 KB tile is covered exactly once
    A = np.array([[tile[sw128_offset(r, k // 4) // 4 + k % 4] for k in range(KT)] for r in range(KP)])
    np.testing.assert_array_equal(A, Y.T)  # row = feature (M / N index), K = neighbour
    for ks in range(KT // 8):  # one MMA reads 8 neighbours: the descriptor start moves by 32 bytes
        sub = np.array([[tile[sw128_offset(r, (8 * ks + kk) // 4) // 4 + (8 * ks + kk) % 4] for kk in range(8)]
                        for r in range(KP)])
        np.testing.assert_array_equal(sub, Y[8 * ks: 8 * ks + 8].T)


def test_kmajor_stores_are_bank_conflict_free():
    for pw in range(4):
        for j in range(4):
            for half in range(2):
                for quarter in range(4):  # an STS.128 is served a quarter warp (128 bytes) at a time
                    slots = {(sw128_offset(l + 32 * j, 2 * pw + half) % 128) // 16
                             for l in range(8 * quarter, 8 * quarter + 8)}
                    assert len(slots) == 8


def test_fused_epilogue_symmetrisation_visits_every_pair_once():
    seen = {}
    for row in range(KP):
        for k in range(1, KP // 2 + 1):
            if k == KP // 2 and row >= KP // 2:
                break
            c = (row + k) & (KP - 1)
            key = (min(row, c), max(row, c))
            seen[key] = seen.get(key, 0) + 1
    assert len(seen) == KP * (KP - 1) // 2 and set(seen.values()) == {1}
