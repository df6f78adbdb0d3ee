TSR_CONV_VERBOSE=1 timeout 600 python tools/microbench_cluster.py 2> gpurun_out/cluster_verbose.log; grep -c "cluster=2" gpurun_out/cluster_verbose.log; grep "cluster=2" gpurun_out/cluster_verbose.log | sort | uniq -c | head -5
python -c "
import sys; sys.path.insert(0,'.')
from torchsr_b200 import ops; ops.check_watchdog(); print('watchdog ok')"
