import sys
sys.path.insert(0,'.')
import ba_b200
syn = ba_b200.synthetic
seq = syn.make_tum_sequence(800, 60000, 400000, seed=3)
p = syn.window_problem(seq, 0, 799).problem
s = ba_b200.GpuSolver(max_num_iterations=2, function_tolerance=0.0, parameter_tolerance=0.0, gradient_tolerance=0.0)
s.upload(p)
summ = s.solve()
print(summ.solve_ms, summ.num_iterations)
