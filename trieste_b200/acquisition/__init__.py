from .function import (  # noqa: F401
    BatchMonteCarloExpectedImprovement,
    ExpectedImprovement,
    LogExpectedImprovement,
    NegativeLowerConfidenceBound,
    batch_monte_carlo_expected_improvement,
    expected_improvement,
    log_expected_improvement,
    lower_confidence_bound,
)
from .interface import AcquisitionFunctionBuilder, SingleModelAcquisitionBuilder  # noqa: F401
