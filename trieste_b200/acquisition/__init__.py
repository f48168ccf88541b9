from .function import (  # noqa: F401
    AugmentedExpectedImprovement,
    augmented_expected_improvement,
    BatchMonteCarloExpectedImprovement,
    ExpectedImprovement,
    LogExpectedImprovement,
    NegativeLowerConfidenceBound,
    ProbabilityOfFeasibility,
    ProbabilityOfImprovement,
    batch_monte_carlo_expected_improvement,
    expected_improvement,
    log_expected_improvement,
    lower_confidence_bound,
    probability_below_threshold,
)
from .interface import AcquisitionFunctionBuilder, SingleModelAcquisitionBuilder  # noqa: F401
