"""The extra legs of bench.py (fused-block table, WanVAE decode leg) run only on a GPU box and sit behind a
try / except in main(): this test makes a failure in one of them loud in the driver's GPU test tier."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_bench_extra_legs_run():
    import bench
    dev = torch.device("cuda:0")
    burst, sustained, _, _ = bench.peaks()
    rows = bench.block_table(dev, burst, sustained, shapes=bench.BLOCK_SHAPES[1:3])      # 4 x 1560 and 1 x 1560
    assert len(rows) == 2
    for r in rows:
        assert r["block_ms"] > 0 and 0.05 < r["frac_of_burst"] < 1.0, r
    v = bench.vae_leg(dev, burst)
    assert v["finite"] and v["ms"] > 0 and 0.05 < v["frac_of_burst"] < 1.0, v
