"""pipe_kernel's opt-in instantiations (DFGPU_PIPE_VAR bits: 1 = L2 prefetch of the survivors' argument sectors at the start of
phase B, 2 = of the tile's key column at the start of phase A, 8 = lane-paired REDs in the aggregate sink, 32 = 256-bit column loads; 43 is the default, 0 the kernel as first measured): they move data earlier or regroup the same atomics, so the fused Q3-shaped plans must produce exactly what
the default instantiation, the oracle's unfused operator chain and pandas produce."""
import pytest

import test_gpu_pipeline as TP
import test_gpu_q3_device_pipeline as TQ

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("var", ["0", "1", "2", "3", "8", "9", "11", "43"])
def test_pipeline_variants_match_the_oracle_chain(gpu_ctx, monkeypatch, var):
    monkeypatch.setenv("DFGPU_PIPE_VAR", var)
    TP.test_pipeline_q3_shape_matches_unfused_oracle_chain(gpu_ctx, False, True, None)       # integer fast evaluator, four aggregates
    TP.test_pipeline_q3_shape_matches_unfused_oracle_chain(gpu_ctx, True, True, 33_333)      # NULL keys / arguments, several batches
    TQ.test_q3_fused_pipelines_match_pandas_and_unfused(gpu_ctx, 0.3)                        # one SUM: the paired sink's shape
