set -e
cd $GRAFT_REPO_ROOT
run() { python -c "
import importlib, sys
sys.path.insert(0,'.')
L = importlib.import_module('bayesian_inference_with_explicit_and_implicit_prior_knowledge_b200._lib'); L.build(force=True)" ; python tools/ticks_bench.py 3 2>&1 | grep -v Warn | head -9; PGAS_SPLIT_SERIAL=1 python tools/ticks_bench.py 3 2>&1 | grep -v Warn | head -2; }
echo "== 512x4"; PGAS_NVCC_EXTRA="" run
echo "== 256x8"; PGAS_NVCC_EXTRA="-DPGAS_WK_NT=256 -DPGAS_WK_PPT=8" run
echo "== 1024x2"; PGAS_NVCC_EXTRA="-DPGAS_WK_NT=1024 -DPGAS_WK_PPT=2" run
