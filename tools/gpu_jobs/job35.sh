#!/bin/bash
# large parity run on the round's last commit (every scoring kernel form, fused similarity, statistics plain / hinted / selected)
timeout 400 python tests/large_parity_check.py > gpurun_out/r2p_large_parity.json 2> gpurun_out/r2p_large_parity.err; echo "large parity rc=$?"
python3 -c "
import json
d=json.loads(open('gpurun_out/r2p_large_parity.json').read().strip().splitlines()[-1])
print({k:v for k,v in d.items() if 'mismatch' in k or k=='ok'})"
