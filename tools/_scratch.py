import os, sys, time
os.environ['MDB200_TIMING'] = '1'
sys.path.insert(0, os.getcwd())
import numpy as np
import workloads
import matrix_decomposition_b200 as matrix
psd = matrix.approximation.positive_semidefinite
for n in (3000, 4000, 4096, 8192):
    A, kw = workloads.shifted_goe(n, 5, 1.05)
    kw = dict(kw, permutation='decreasing_diagonal_values')
    for rep in range(3):
        t = time.perf_counter(); dec = psd.decomposition(A, **kw); t1 = time.perf_counter(); L = dec.L; t2 = time.perf_counter()
        del dec, L
        t3 = time.perf_counter()
        print('n', n, 'decomposition %.2f ms  .L %.2f ms  del %.2f ms' % ((t1 - t) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3), flush=True)
