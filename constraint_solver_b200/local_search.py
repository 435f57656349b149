"""Host-side mirror of local-search/src/local_search.rs `LocalSearch` (:253-343) for the two
plug-ins the reference ships.

Same constructor arguments and `execute(start, allow_no_improvement_for)` shape as the
reference; the neighbourhood enumeration, delta scoring, best-move selection and acceptance
all run on the device (cs_*_local_search_one).  `window_size` is honoured when the proposer is
the reference's own sampled one (`reference=True`); otherwise the device evaluates the whole
neighbourhood (window = infinity).  `all_solutions_capacity` / `all_solution_iteration_expiry`
are accepted and cannot change a result: the reference's tabu set is always {current}
(local_search.rs:182-195), which the device implements by skipping identity moves.
"""
from __future__ import annotations

import numpy as np

from .nqueens import (NQueensChains, NQueensMoveProposer, NQueensScore, NQueensSolution,
                      NQueensSolutionScoreCalculator, ScoredSolution)
from .scheduling import ScheduleChains, ScheduleMoveProposer, ScheduleScore


class LocalSearch:
    def __init__(self, move_proposer, solution_score_calculator, max_iterations: int,
                 window_size: int = 0, best_solutions_capacity: int = 16,
                 all_solutions_capacity: int = 10_000, all_solution_iteration_expiry: int = 100_000,
                 rng=None):
        self.move_proposer = move_proposer
        self.solution_score_calculator = solution_score_calculator
        self.max_iterations = max_iterations
        self.window_size = window_size
        self.rng = rng
        if isinstance(move_proposer, NQueensMoveProposer):
            self._kind = "nq"
            self._engine = NQueensChains(move_proposer.board_size, 1,
                                         neighbourhood=move_proposer.neighbourhood)
        elif isinstance(move_proposer, ScheduleMoveProposer):
            self._kind = "es"
            mp = move_proposer
            self._engine = ScheduleChains(mp.n_days, mp.employees, start_weekday=mp.start_weekday,
                                          holidays=mp.holidays, reference_proposer=mp.reference)
            if mp.reference and window_size:
                self._engine.set_window(window_size)
        else:
            raise TypeError("LocalSearch runs on the device for the reference's two plug-ins: pass an "
                            "NQueensMoveProposer or a ScheduleMoveProposer")

    def execute(self, start, allow_no_improvement_for: int):
        """local_search.rs:301-342"""
        if self._kind == "nq":
            best, score = self._engine.local_search_one(start.rows, allow_no_improvement_for,
                                                        self.max_iterations)
            return ScoredSolution(NQueensScore(score), NQueensSolution(best))
        best, hard, soft = self._engine.local_search_one(np.asarray(start), allow_no_improvement_for,
                                                         self.max_iterations)
        return ScheduleScore(float(hard), float(soft)), best


__all__ = ["LocalSearch", "NQueensSolutionScoreCalculator", "ScheduleMoveProposer"]
