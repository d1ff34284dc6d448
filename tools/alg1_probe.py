"""first vs second call of the shipped single-mass-oscillator Algorithm1 (developer aid): where do the seconds of a first call go?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
t0 = time.time()
import torch
torch.cuda.init(); torch.zeros(1, device="cuda"); torch.cuda.synchronize()
t1 = time.time()
import src.SingleMassOscillator as E
import bayesian_inference_with_explicit_and_implicit_prior_knowledge_b200.random as rnd
t2 = time.time()
key = E.key
times = []
for i in range(3):
    key, k = rnd.split(key)
    a = time.time(); out = E.SMO_Algorithm1(k); torch.cuda.synchronize(); times.append(time.time() - a)
import numpy as np
a = time.time(); tr = E.SMO_SSM.tables(np.asarray(E.F_ext, dtype=float), 2, [1]); tt = time.time() - a
print(dict(torch_and_context_s=t1 - t0, import_example_module_s=t2 - t1, algorithm1_calls_s=times, host_tracing_tables_s=tt))
