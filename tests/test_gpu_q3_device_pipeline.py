"""Config C4 device resident (scripts/q3_device_pipeline.py, the pipeline bench.py times at SF100): at small scale factors the
result must equal an independent pandas evaluation of the same generated tables — bit-exact (int64 fixed-point money)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("sf", [0.01, 0.3])
def test_q3_device_pipeline_matches_pandas(gpu_ctx, sf):
    import q3_device_pipeline as Q
    customer, orders, lineitem = Q.gen_tables(gpu_ctx, sf)
    res, stages = Q.run_q3(gpu_ctx, customer, orders, lineitem)
    got = Q.result_rows(gpu_ctx, res)
    exp = Q.q3_expected(customer.host(gpu_ctx), orders.host(gpu_ctx), lineitem.host(gpu_ctx))
    assert stages["groups"] == len(exp) > 0
    assert got == exp
    for b in res:
        b.release()


@pytest.mark.parametrize("sf", [0.01, 0.3])
def test_q3_fused_pipelines_match_pandas_and_unfused(gpu_ctx, sf):
    """the three fused pipelines (dfgpu_pipeline) give the rows of the operator-by-operator path and of pandas, bit-exact"""
    import q3_device_pipeline as Q
    customer, orders, lineitem = Q.gen_tables(gpu_ctx, sf)
    res, stages = Q.run_q3_fused(gpu_ctx, customer, orders, lineitem)
    got = Q.result_rows(gpu_ctx, res)
    ref, rstages = Q.run_q3(gpu_ctx, customer, orders, lineitem)
    assert Q.result_fingerprint(gpu_ctx, res) == Q.result_fingerprint(gpu_ctx, ref)
    for k in ("customer_building", "orders_of_building_customers", "joined_rows", "groups"):
        assert stages[k] == rstages[k], k
    exp = Q.q3_expected(customer.host(gpu_ctx), orders.host(gpu_ctx), lineitem.host(gpu_ctx))
    assert stages["groups"] == len(exp) > 0
    assert got == exp
    for b in res + ref:
        b.release()
