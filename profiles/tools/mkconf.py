import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, robocheck_b200 as rc
g = np.load(os.path.join(rc.REPO_ROOT, "tests/golden/hard_confs10k.npz"))
rc.writeConfigurationToFileTagged(g["confs"][:400], g["tags"][:400], "gpurun_out/small.conf")
