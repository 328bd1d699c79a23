"""pipe_kernel's prefetch instantiations (DFGPU_PIPE_VAR bits: 4 = four survivors per lane and phase-B round; 1 | 2 = L2 prefetch of the survivors' argument sectors at the start
of phase B / of the tile's key column at the start of phase A): they only move data earlier, so the fused Q3-shaped plan must
produce exactly what the default instantiation and the oracle's unfused operator chain produce."""
import pytest

import test_gpu_pipeline as TP

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("var", ["1", "2", "3", "4", "5", "7"])
def test_pipeline_prefetch_variants_match_the_oracle_chain(gpu_ctx, monkeypatch, var):
    monkeypatch.setenv("DFGPU_PIPE_VAR", var)
    TP.test_pipeline_q3_shape_matches_unfused_oracle_chain(gpu_ctx, False, True, None)       # integer fast evaluator
    TP.test_pipeline_q3_shape_matches_unfused_oracle_chain(gpu_ctx, True, True, 33_333)      # NULL keys / arguments, several batches
