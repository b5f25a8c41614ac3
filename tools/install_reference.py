"""Copy the UNMODIFIED reference's Python packages into baseline/_ref/ (git-ignored, NOT
gpurun-ignored: it travels to the GPU box with the snapshot, where /root/reference does not exist).

    python tools/install_reference.py            # /root/reference -> baseline/_ref

The reference has no setup.py / pyproject.toml, so `pip install --target baseline/_ref` has nothing to
build; its importable form is its source tree (hsg/, lib/, pyscripts/, configs/ and the two colour
maps: 0.5 MB).  Used by tests/test_reference_step.py (the reference's own model / gather / loss code,
run unpatched on CUDA and again over hsg_b200.patch()), by tools/run_reference_train.py (the
reference's unchanged pyscripts/train/train.py on synthetic data) and by `bench.py --impl reference`
(the reference's own functions on the host cores).  Nothing under hsg_b200/ imports it.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, 'baseline', '_ref')
PARTS = ['hsg', 'lib', 'pyscripts', 'configs', 'misc/colormapvoc.mat', 'misc/colormapcs.mat', 'LICENSE']


def install(src='/root/reference', dst=DST, force=False):
  """Returns the install directory, or None when there is neither a source tree nor an earlier copy."""
  marker = os.path.join(dst, '.installed_from')
  if not os.path.isdir(src):
    return dst if os.path.exists(marker) else None
  if os.path.exists(marker) and not force:
    return dst
  if os.path.isdir(dst):
    shutil.rmtree(dst)
  for part in PARTS:
    s, d = os.path.join(src, part), os.path.join(dst, part)
    if os.path.isdir(s):
      shutil.copytree(s, d, ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    elif os.path.exists(s):
      os.makedirs(os.path.dirname(d), exist_ok=True)
      shutil.copy(s, d)
  for root, dirs, files in os.walk(dst):          # the source tree is read-only; the copy must be removable
    for n in dirs + files:
      os.chmod(os.path.join(root, n), 0o755 if n in dirs else 0o644)
  with open(marker, 'w') as f:
    f.write(src + '\n')
  return dst


if __name__ == '__main__':
  print(install(force='--force' in sys.argv))
