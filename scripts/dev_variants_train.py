"""Dev helper (GPU box): train-step time (P=4, N=50k, 32 bags) for each prebuilt library variant."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for lib in sorted(glob.glob(os.path.join(ROOT, "vlsa_b200/lib/variants/*.so"))):
    print("==", os.path.basename(lib), flush=True)
    env = dict(os.environ, VLSA_B200_LIB=lib, DEV_QUICK="1")
    subprocess.run([sys.executable, os.path.join(ROOT, "scripts/dev_train_time.py"), "child"], env=env)
