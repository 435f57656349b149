"""Host-side mirror of local-search/src/local_search.rs `LocalSearch` (:253-343).

Same constructor arguments and `execute(start, allow_no_improvement_for)` shape as the
reference; the neighbourhood enumeration, delta scoring, best-move selection and acceptance
all run on the device (cs_*_local_search_one).  `window_size` is accepted for signature
compatibility: the device always evaluates the whole neighbourhood (window = infinity).
"""
from __future__ import annotations

from .nqueens import (NQueensChains, NQueensMoveProposer, NQueensScore, NQueensSolution,
                      NQueensSolutionScoreCalculator, ScoredSolution)


class LocalSearch:
    def __init__(self, move_proposer, solution_score_calculator, max_iterations: int,
                 window_size: int = 0, best_solutions_capacity: int = 16,
                 all_solutions_capacity: int = 10_000, all_solution_iteration_expiry: int = 100_000,
                 rng=None):
        if not isinstance(move_proposer, NQueensMoveProposer):
            raise TypeError("this LocalSearch is the n-queens device drop-in; "
                            "scheduling uses constraint_solver_b200.scheduling.LocalSearch")
        self.move_proposer = move_proposer
        self.solution_score_calculator = solution_score_calculator
        self.max_iterations = max_iterations
        self.window_size = window_size
        self.rng = rng
        self._engine = NQueensChains(move_proposer.board_size, 1,
                                     neighbourhood=move_proposer.neighbourhood)

    def execute(self, start: NQueensSolution, allow_no_improvement_for: int) -> ScoredSolution:
        """local_search.rs:301-342"""
        best, score = self._engine.local_search_one(start.rows, allow_no_improvement_for,
                                                    self.max_iterations)
        return ScoredSolution(NQueensScore(score), NQueensSolution(best))


__all__ = ["LocalSearch", "NQueensSolutionScoreCalculator"]
