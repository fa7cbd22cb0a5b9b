"""Transcribes the block-sparse known-answer problems of the reference's modified Ceres
(CERES/internal/ceres/linear_least_squares_problems.cc: problem 2 :283-404 with the hand-computed
values of the comment at :135-178, problem 3 :406-515, problem 4 :517-625) into
tests/golden/ceres_llsq_problems.json.  Data only (matrix entries and expected numbers), written
out by hand from the cited lines; run from the repo root: python tests/golden/make_ceres_fixtures.py"""
import json
import os

problems = {
    "problem2": dict(
        col_sizes=[1, 1, 1, 1, 1], num_eliminate_blocks=2,
        row_sizes=[1, 1, 1, 1, 1, 1],
        row_ptr=[0, 2, 4, 6, 8, 10, 13],
        cell_col=[0, 2, 0, 3, 1, 4, 1, 2, 1, 2, 2, 3, 4],
        values=[1, 2, 3, 4, 5, 6, 7, 8, 9, 1, 1, 1, 1],
        b=[0, 1, 2, 3, 4, 5], D=[1, 1, 1, 1, 1],
        golden=dict(
            c=[3, 67, 33, 9, 17],
            AtA=[[10, 0, 2, 12, 0], [0, 155, 65, 0, 30], [2, 65, 70, 1, 1], [12, 0, 1, 17, 1],
                 [0, 30, 1, 1, 37]],
            S=[[42.3419, -1.4000, -11.5806], [-1.4000, 2.6000, 1.0000], [-11.5806, 1.0000, 31.1935]],
            r=[4.3032, 5.4000, 5.0323],
            S_solve_r=[0.2102, 2.1367, 0.1388],
            A_solve_b=[-2.3061, 0.3172, 0.2102, 2.1367, 0.1388])),
    "problem3": dict(
        col_sizes=[1, 1], num_eliminate_blocks=2,
        row_sizes=[1, 1, 1, 1, 1], row_ptr=[0, 1, 2, 3, 4, 5], cell_col=[0, 0, 1, 1, 1],
        values=[1, 3, 5, 7, 9], b=[0, 1, 2, 3, 4], D=[1, 1]),
    "problem4": dict(
        col_sizes=[2, 3, 2], num_eliminate_blocks=1,
        row_sizes=[2, 1], row_ptr=[0, 2, 4], cell_col=[0, 2, 1, 2],
        values=[1, 2, 1, 4, 1, 1, 5, 6, 9, 0, 0, 3, 1],
        b=[0, 1, 2], D=[100, 200, 300, 400, 500, 600, 700]),
}
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ceres_llsq_problems.json")
with open(out, "w") as f:
    json.dump(problems, f, indent=1)
print("wrote", out)
