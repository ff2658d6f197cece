"""numpy model of the link kernel's work decomposition (csrc/shard.cu: launch_push + peer_push_bulk_kernel): the host-side
choice of segment size / rows per item / contiguous fast path and the device-side item -> (job, rows, column offset) decode,
replayed over numpy buffers for the geometries the row-sharded path produces (contiguous 16 MB blocks) and for pitched and tiny
ones.  Every destination element must be written exactly once with the right source value, whatever the number of CTAs."""
import numpy as np
import pytest

STAGE_BYTES = 32768


def plan_push(row_elems, rows, src_pitch, dst_pitch):
    """launch_push(): returns (seg_bytes, segs_per_row, rows_per_item, items_per_job, contig)."""
    row_bytes = 8 * row_elems
    seg_bytes = min(row_bytes, STAGE_BYTES)
    segs_per_row = row_bytes // seg_bytes
    rows_per_item = STAGE_BYTES // seg_bytes
    contig = segs_per_row == 1 and src_pitch == row_elems and dst_pitch == row_elems
    if not contig and rows_per_item > 64:
        rows_per_item = 64
    items_per_job = -(-rows // rows_per_item) * segs_per_row
    return seg_bytes, segs_per_row, rows_per_item, items_per_job, contig


def run_push(jobs, row_elems, rows, src_pitch, dst_pitch, ctas):
    """jobs: list of (src array, dst array) of complex64, 1-D views with the given pitches.  Emulates every CTA's loop."""
    seg_bytes, segs_per_row, rows_per_item, items_per_job, contig = plan_push(row_elems, rows, src_pitch, dst_pitch)
    nbig = len(jobs)
    total = items_per_job * nbig
    grid = max(1, min(ctas, total))
    writes = [np.zeros(d.size, np.int32) for _, d in jobs]
    for cta in range(grid):
        k = 0
        while cta + k * grid < total:
            idx = cta + k * grid
            j, w = idx % nbig, idx // nbig
            rg = w // segs_per_row
            c0 = (w - rg * segs_per_row) * (seg_bytes // 8)
            row0 = rg * rows_per_item
            nrows = min(rows_per_item, rows - row0)
            assert nrows * seg_bytes <= STAGE_BYTES
            src, dst = jobs[j]
            if contig:
                n = nrows * seg_bytes // 8
                dst[row0 * dst_pitch:row0 * dst_pitch + n] = src[row0 * src_pitch:row0 * src_pitch + n]
                writes[j][row0 * dst_pitch:row0 * dst_pitch + n] += 1
            else:
                for r in range(nrows):
                    so, do, n = (row0 + r) * src_pitch + c0, (row0 + r) * dst_pitch + c0, seg_bytes // 8
                    dst[do:do + n] = src[so:so + n]
                    writes[j][do:do + n] += 1
            k += 1
    return writes


@pytest.mark.parametrize("row_elems,rows,src_pitch,dst_pitch,peers,ctas", [
    (1024, 2048, 1024, 1024, 7, 24),      # 16384^2 over 8 GPUs, half planes: contiguous 16 MB blocks
    (1024, 1700, 1024, 1024, 7, 16),      # image height not a multiple of the item size
    (4096, 512, 4096, 4096, 1, 32),       # 32 KB rows: one segment per row
    (8192, 64, 8192, 8192, 3, 5),         # 64 KB rows: two segments per row
    (1024, 96, 8192, 1024, 7, 24),        # pitched source (row-interleaved staging, the first layout tried)
    (1024, 96, 1024, 8192, 7, 24),        # pitched destination
    (16, 33, 16, 16, 3, 3),               # tiny rows (tests): many rows per item
    (2, 5, 64, 2, 1, 4),                  # 16-byte rows, pitched source
])
def test_push_decode_covers_every_element_once(row_elems, rows, src_pitch, dst_pitch, peers, ctas):
    rng = np.random.default_rng(row_elems * 31 + rows)
    jobs = []
    for _ in range(peers):
        src = (rng.standard_normal(rows * src_pitch) + 1j * rng.standard_normal(rows * src_pitch)).astype(np.complex64)
        dst = np.zeros(rows * dst_pitch, np.complex64)
        jobs.append((src, dst))
    writes = run_push(jobs, row_elems, rows, src_pitch, dst_pitch, ctas)
    for (src, dst), wr in zip(jobs, writes):
        s2 = src.reshape(rows, src_pitch)[:, :row_elems]
        d2 = dst.reshape(rows, dst_pitch)
        assert np.array_equal(d2[:, :row_elems], s2)
        w2 = wr.reshape(rows, dst_pitch)
        assert (w2[:, :row_elems] == 1).all() and (w2[:, row_elems:] == 0).all()
