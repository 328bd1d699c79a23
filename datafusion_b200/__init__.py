"""datafusion_b200 — B200 (sm_100a) kernel layer for DataFusion's FilterExec / HashJoinExec /
AggregateExec hot paths.  See DESIGN.md.  The CUDA library is mandatory: no CPU fallback."""
from . import capi  # noqa: F401

__all__ = ["capi"]
