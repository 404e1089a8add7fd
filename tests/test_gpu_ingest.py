"""GPU: Decimal128 ingest with host-side narrowing (host_pool.hpp / import_batch) is bit-exact: chunks whose
values fit int32 travel as 4 bytes, int64-range chunks as 8, genuinely wide chunks unchanged; negative values,
chunk boundaries and Arrow array offsets included.  Checked by exporting the registered table back."""
import numpy as np
import pyarrow as pa
import pytest

import ballista_b200 as bb

pytestmark = pytest.mark.gpu


def _dec_col(lo: np.ndarray, hi: np.ndarray, typ):
    pairs = np.empty(2 * lo.size, dtype=np.int64)
    pairs[0::2] = lo
    pairs[1::2] = hi
    return pa.Array.from_buffers(typ, lo.size, [None, pa.py_buffer(pairs.tobytes())])


def test_narrowed_ingest_round_trip(gpu):
    n = 5_000_000  # more than one 4M-row chunk
    rng = np.random.default_rng(5)
    typ = pa.decimal128(38, 2)
    small = rng.integers(-2_000_000_000, 2_000_000_000, n, dtype=np.int64)
    a_lo, a_hi = small.copy(), small >> 63
    b_lo = small.copy()
    b_lo[4_500_000] = 9_000_000_000_000          # second chunk needs int64
    b_lo[4_500_001] = -9_000_000_000_000
    b_hi = b_lo >> 63
    c_lo, c_hi = small.copy(), small >> 63
    c_hi[123_456] = 77                            # first chunk holds a value beyond 64 bits
    c_lo[123_457], c_hi[123_457] = 5, -3
    batch = pa.record_batch([_dec_col(a_lo, a_hi, typ), _dec_col(b_lo, b_hi, typ), _dec_col(c_lo, c_hi, typ),
                             pa.array(rng.integers(0, 1 << 40, n), type=pa.int64())], names=["a", "b", "c", "k"])
    gpu.set_config("b200.ingest.narrow_decimals", "on")
    try:
        saved0 = gpu.counter("ingest_bytes_saved")
        for name, bt in (("ing", batch), ("ing_off", batch.slice(7, 1_234_567))):  # sliced: non-zero Arrow offset
            gpu.drop_table(name)
            gpu.register_batch(name, 0, bt)
            got = gpu.export_table(name, 0)
            assert got.num_rows == bt.num_rows
            for i in range(bt.num_columns):
                assert got.column(i).equals(bt.column(i)), (name, bt.schema.names[i])
            gpu.drop_table(name)
        # a: 12 B/row saved everywhere; b: 12 B in chunk 1, 8 B in chunk 2; c: nothing in chunk 1, 12 B in chunk 2
        assert gpu.counter("ingest_bytes_saved") - saved0 >= 12 * n
    finally:
        gpu.set_config("b200.ingest.narrow_decimals", "auto")
