"""Host-side model of the streamed kernel's converter indexing (t-vq-vae-trajgen_b200/csrc/tvq_fwd_stream.cuh, the
`STAGED` branch of the converter warps): a lane owns a row (d <= 64) or half a row (d <= 128) of a staged 8 KB block
and walks its 16 chunks in an order rotated by the row.  Checked here, without a GPU: every (row, chunk) of a block is
converted exactly once, the destination is the UMMA K-major SWIZZLE_128B address of that element group, and the shared-
memory accesses stay within the bank-conflict budget the design states (LDS.128 at most 2-way, STS.64 conflict-free).
The GPU parity tests (tests/test_parity_gpu.py::test_stream_path_*) check the same code against the oracle."""
import pytest

A_SLAB = 128 * 128          # bytes of one A slab: 128 rows x 128 bytes (64 bf16)


def lane_plan(dp, lane, j):
    """(row in block, fp32 chunk index cc, source byte offset in the block, destination byte offset in the A tile
    relative to the block's first row) of step j (0..15) of a lane — the formulas of the kernel."""
    f = dp // 4                         # 16-byte fp32 chunks per padded row
    xrb = 512 // f                      # rows per staged block
    lpr = 32 // xrb                     # lanes per row
    brw, sub = lane // lpr, lane % lpr
    rot = brw + 8 * sub
    ccl = (j + rot) & 15
    cc = sub * 16 + ccl
    rowbytes = dp * 4
    src = brw * rowbytes + sub * 256 + ccl * 16
    rx = brw & 7
    dst = sub * A_SLAB + brw * 128 + (((ccl >> 1) ^ rx) << 4) + ((ccl & 1) << 3)
    return brw, cc, src, dst


def sw128_address(row, col):
    """Byte address of bf16 element (row, col) of a K-major SWIZZLE_128B tile cut into 64-column slabs."""
    slab, c = col // 64, col % 64
    chunk16 = (c * 2) // 16             # 16-byte chunk within the row's 128 bytes
    return slab * A_SLAB + row * 128 + ((chunk16 ^ (row & 7)) << 4) + (c * 2) % 16


@pytest.mark.parametrize("dp", [64, 128])
def test_every_chunk_once_and_at_its_swizzled_address(dp):
    f, xrb = dp // 4, 512 // (dp // 4)
    seen = {}
    for lane in range(32):
        for j in range(16):
            row, cc, src, dst = lane_plan(dp, lane, j)
            assert 0 <= row < xrb and 0 <= cc < f
            assert (row, cc) not in seen
            seen[(row, cc)] = (src, dst)
            assert src == row * dp * 4 + cc * 16                     # the fp32 chunk of that row in the staged block
            assert dst == sw128_address(row, 4 * cc)                 # four bf16 = 8 bytes from element column 4 cc
    assert len(seen) == xrb * f


def _wavefronts(addrs, width, group):
    """Shared-memory wavefronts of one warp-wide access: lanes are served `group` at a time, a wavefront moves one
    `width`-byte word per 128-byte bank row position."""
    total = 0
    for g0 in range(0, 32, group):
        per_bank = {}
        for a in addrs[g0:g0 + group]:
            per_bank.setdefault((a // width) % (128 // width), set()).add(a)
        total += max(len(v) for v in per_bank.values())
    return total


@pytest.mark.parametrize("dp,d", [(64, 64), (64, 48), (128, 128), (128, 100)])
def test_bank_conflicts_within_budget(dp, d):
    for j in range(16):
        src, dst = [], []
        for lane in range(32):
            row, cc, _, dd = lane_plan(dp, lane, j)
            lpr = 32 // (512 // (dp // 4))
            src.append(row * d * 4 + (lane % lpr) * 256 + (cc % 16) * 16)      # rows are d * 4 bytes apart in the block
            dst.append(dd)
        assert _wavefronts(src, 16, 8) <= 8        # LDS.128: 4 wavefronts when conflict-free, at most 2-way
        assert _wavefronts(dst, 8, 16) == 2        # STS.64: conflict-free
