"""Mirror of the reference's option structs (``src/ilqr_options.hh:4-22``).

Field names and meaning are the reference's; ``ConvergenceCriteria.max_iters`` is a
float because it is a ``double`` there (``ilqr_options.hh:14``).  The three extra
fields of :class:`ILQROptions` are batch-only extensions whose defaults reproduce the
reference exactly.
"""
from __future__ import annotations

from dataclasses import dataclass, field


@dataclass
class LineSearchParams:  # ilqr_options.hh:4-8
    step_update: float = 0.5
    desired_reduction_frac: float = 0.5
    max_iters: int = 100


@dataclass
class ConvergenceCriteria:  # ilqr_options.hh:11-15
    rtol: float = 1e-12
    atol: float = 1e-12
    max_iters: float = 100.0


@dataclass
class ILQROptions:  # ilqr_options.hh:18-22
    line_search_params: LineSearchParams = field(default_factory=LineSearchParams)
    convergence_criteria: ConvergenceCriteria = field(default_factory=ConvergenceCriteria)
    populate_debug: bool = False
    # extensions (defaults = reference behaviour)
    symmetrize_vxx: bool = False
    num_parallel_alphas: int = 1
    quu_regularization: float = 0.0

    def __eq__(self, other):  # operator== (ilqr_options.cc:17-21), without its `&&` slip
        return (isinstance(other, ILQROptions)
                and self.line_search_params == other.line_search_params
                and self.convergence_criteria == other.convergence_criteria
                and self.populate_debug == other.populate_debug
                and self.symmetrize_vxx == other.symmetrize_vxx
                and self.num_parallel_alphas == other.num_parallel_alphas
                and self.quu_regularization == other.quu_regularization)
