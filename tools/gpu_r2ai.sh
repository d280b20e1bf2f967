#!/bin/bash
# round-2 call AI: last sanity of the final in-tree build (smoke + the golden-fixture model test)
O=gpurun_out
timeout 70 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2ai_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2ai_smoke.log
timeout 60 python -m pytest tests/test_gpu_model.py -m gpu -q -k "golden or uint8" > $O/r2ai_pytest.log 2>&1; echo "rc=$?" >> $O/r2ai_pytest.log
tail -2 $O/r2ai_smoke.log; tail -2 $O/r2ai_pytest.log
