"""TEST INFRASTRUCTURE (checker only, never the product path): stage the UNMODIFIED reference as `oracle/_ref`.

    python oracle/stage_ref.py            # build container: needs /root/reference

The reference (neherlab/treetime 0.12.1) is pure Python.  It is installed with pip, untouched, from a
scratch copy of /root/reference into the git-ignored directory oracle/_ref, which -- unlike /root/reference
-- travels to the GPU box with the gpurun snapshot.  There the `-m gpu` drop-in tests run the real
`treetime.TreeAnc` / `treetime.TreeTime` next to the accelerated classes, and `bench.py --impl reference`
times the real `TreeAnc.infer_ancestral_sequences(marginal=True)`.  Biopython is not in this image:
oracle/bioshim (container stubs without numerics) stands in for it (see oracle/refenv.py).
No reference source enters the repository history (.gitignore lists oracle/_ref/).
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, '_ref')
SRC = os.environ.get('TREETIME_REFERENCE', '/root/reference')


def staged():
    return os.path.isfile(os.path.join(DEST, 'treetime', 'treeanc.py'))


def stage(force=False, verbose=False):
    """Returns 'present' / 'installed' / 'unavailable: why'."""
    if staged() and not force:
        return 'present'
    if not os.path.isdir(os.path.join(SRC, 'treetime')):
        return 'unavailable: %s not found' % SRC
    tmp = tempfile.mkdtemp(prefix='ttref_')
    try:
        work = os.path.join(tmp, 'reference')
        shutil.copytree(SRC, work, ignore=shutil.ignore_patterns('.git', 'docs', 'benchmarking', '__pycache__'))   # the source tree is read-only
        if os.path.isdir(DEST):
            shutil.rmtree(DEST)
        cmd = [sys.executable, '-m', 'pip', 'install', '--no-index', '--no-build-isolation', '--no-deps', '--find-links',
               '/opt/wheelhouse', '--target', DEST, work]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose:
            sys.stderr.write(r.stdout[-2000:] + r.stderr[-2000:])
        if r.returncode != 0 or not staged():
            return 'unavailable: pip install failed: %s' % (r.stderr.strip().splitlines() or ['?'])[-1]
        return 'installed'
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == '__main__':
    print(stage(force='--force' in sys.argv, verbose=True))
