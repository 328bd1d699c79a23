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


@pytest.mark.parametrize("sf", [0.05])
def test_q3_fused_with_decimal128_money_equals_int64_money(gpu_ctx, sf):
    """the same tables with l_extendedprice / l_discount as Decimal128(15,2) (the reference's TPC-H schema): the Decimal128(38,4) sums
    carry the same unscaled integers as the int64 fixed-point run, so rows, keys and sums must agree one for one"""
    import numpy as np
    import q3_device_pipeline as Q
    from datafusion_b200 import capi as D
    customer, orders, lineitem = Q.gen_tables(gpu_ctx, sf)
    dl = Q.decimal_money(gpu_ctx, lineitem)
    assert dl.types[1] == dl.types[2] == D.decimal128(15, 2)
    res, stages = Q.run_q3_fused(gpu_ctx, customer, orders, dl)
    ref, rstages = Q.run_q3_fused(gpu_ctx, customer, orders, lineitem)
    assert stages == {**rstages, "lookup_bytes": stages["lookup_bytes"]}          # one more accumulator word per record
    assert res[0].column(3).type == D.decimal128(38, 4)
    assert Q.result_fingerprint(gpu_ctx, res) == Q.result_fingerprint(gpu_ctx, ref)
    want = Q.result_rows(gpu_ctx, ref)
    got = []
    for b in res:
        keys = [gpu_ctx.to_host(b.column(i).values, b.num_rows * D.WIDTH[b.column(i).type]).view(D.NP_OF_TYPE[b.column(i).type]).tolist() for i in range(3)]
        words = gpu_ctx.to_host(b.column(3).values, b.num_rows * 16).view(np.uint64).reshape(-1, 2)
        got += list(zip(keys[0], keys[1], keys[2], D.words_to_decimal(words)))
    assert sorted(got) == want and len(want) > 1000
    for b in res + ref:
        b.release()
