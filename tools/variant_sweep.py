#!/usr/bin/env python
"""Time every kernel-variant library under build/variants/ (experiments)."""
import glob, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for lib in sorted(glob.glob(os.path.join(root, "build", "variants", "*.so"))):
    env = dict(os.environ, RPO_B200_LIB=lib)
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "quick_gpu.py")], env=env, capture_output=True, text=True)
    lines = [l for l in out.stdout.splitlines() if l.startswith(("B ", "projected", "project without"))]
    print(os.path.basename(lib), "|", " | ".join(lines) if lines else out.stderr[-400:], flush=True)
