#!/usr/bin/env python
"""Build the UNMODIFIED reference (qutip 5.4.0.dev) out of tree into ``oracle/_ref/``.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package ``qutip_b200`` imports
this directory; it exists so that tests / bench ``--impl reference`` can time and
check against the reference's own CPU implementation of the hot path.

The reference's build system is meson-python, which is not installed here, so the
module list of ``/root/reference/qutip/meson.build:25-64`` is restated as plain
setuptools/Cython extensions (recipe from SURVEY.md section 8c / appendix B).

Outputs go ONLY into ``oracle/_ref/`` (git-ignored, but shipped to the GPU box):
    oracle/_ref/qutip/...                       python sources + built .so
    oracle/_ref/qutip-5.4.0.dev0.dist-info/     metadata stub (importlib.metadata)

Usage:  python oracle/build_ref.py [--jobs N] [--force]
"""
import argparse
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
OUT = os.path.join(HERE, "_ref")

SETUP_PY = r'''
import numpy, os
from setuptools import setup, Extension
from Cython.Build import cythonize
mods = {
 '.': ['_distributions'],
 'core': ['_brtools', '_brtensor'],
 'core/cy': ['_element', 'coefficient', 'lindblad_matrix_form', 'math', 'qobjevo'],
 'core/data': ['add', 'adjoint', 'base', 'block_operations', 'convert', 'csr',
               'data_iterator', 'dense', 'dia', 'dispatch', 'expect', 'inner', 'kron',
               '_local_matmul', 'mean', 'mul', 'norm', 'ode', 'permute', 'pow',
               'project', 'properties', 'ptrace', 'reshape', 'tidyup', 'trace'],
 'piqs': ['_piqs'],
 'solver/cy': ['dysolve', 'nm_mcsolve'],
 'solver/integrator': ['explicit_rk', '_rhs'],
 'solver/sode': ['_sode', 'ssystem'],
}
inc = [numpy.get_include(), 'qutip/core/data']
args = ['-O3', '-funroll-loops', '-std=c++17', '-w']
mac = [('NPY_NO_DEPRECATED_API', 'NPY_1_7_API_VERSION')]
exts = []
for d, names in mods.items():
    for n in names:
        path = os.path.normpath(os.path.join('qutip', d, n))
        exts.append(Extension(path.replace('/', '.'), [path + '.pyx'], include_dirs=inc,
                              extra_compile_args=args, language='c++', define_macros=mac))
exts.append(Extension('qutip.core.data.matmul',
                      ['qutip/core/data/matmul.pyx',
                       'qutip/core/data/src/matmul_csr_vector.cpp',
                       'qutip/core/data/src/matmul_csr_dense.cpp',
                       'qutip/core/data/src/matmul_diag_vector.cpp'],
                      include_dirs=inc, extra_compile_args=args, language='c++',
                      define_macros=mac))
setup(name='qutip_ref_build',
      ext_modules=cythonize(exts, language_level=3, nthreads=JOBS),
      script_args=['build_ext', '--inplace', '-j', str(JOBS)])
'''


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=os.cpu_count() or 4)
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()

    marker = os.path.join(OUT, ".built")
    if os.path.exists(marker) and not a.force:
        print("oracle/_ref already built (use --force to rebuild)")
        return 0
    if not os.path.isdir(os.path.join(REF_SRC, "qutip")):
        print("reference tree not present; cannot build oracle/_ref", file=sys.stderr)
        return 1
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    pkg = os.path.join(OUT, "qutip")
    shutil.copytree(os.path.join(REF_SRC, "qutip"), pkg,
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for root, dirs, files in os.walk(OUT):
        os.chmod(root, 0o755)
        for f in files:
            os.chmod(os.path.join(root, f), 0o644)
    # what qutip/meson.build:5-17 generates
    with open(os.path.join(pkg, "core", "data", "src", "intdtype.h"), "w") as f:
        f.write("typedef int32_t idxint;\nstatic const int _idxint_size = 32;\n")
    version = open(os.path.join(REF_SRC, "VERSION")).read().strip()
    if version.endswith(".dev"):
        version += "0"
    with open(os.path.join(pkg, "version.py"), "w") as f:
        f.write("short_version = %r\nversion = %r\nrelease = False\n"
                % (version.split(".dev")[0], version))
    di = os.path.join(OUT, "qutip-%s.dist-info" % version)
    os.makedirs(di)
    with open(os.path.join(di, "METADATA"), "w") as f:
        f.write("Metadata-Version: 2.1\nName: qutip\nVersion: %s\n" % version)
    with open(os.path.join(di, "RECORD"), "w") as f:
        f.write("")
    with open(os.path.join(di, "INSTALLER"), "w") as f:
        f.write("oracle/build_ref.py\n")
    setup_path = os.path.join(OUT, "_setup_ref.py")
    with open(setup_path, "w") as f:
        f.write("JOBS = %d\n" % a.jobs + SETUP_PY)
    env = dict(os.environ)
    r = subprocess.run([sys.executable, setup_path], cwd=OUT, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        print(r.stdout[-6000:])
        return r.returncode
    # trim: generated C++ and build tree are not needed at run time
    shutil.rmtree(os.path.join(OUT, "build"), ignore_errors=True)
    for root, dirs, files in os.walk(pkg):
        for f in files:
            p = os.path.join(root, f)
            if f.endswith(".pyx"):
                gen = p[:-4] + ".cpp"
                if os.path.exists(gen):
                    os.remove(gen)
            if f.endswith(".so"):
                subprocess.run(["strip", "--strip-unneeded", p], check=False)
    open(marker, "w").write(version + "\n")
    print("built reference %s into %s" % (version, OUT))
    return 0


if __name__ == "__main__":
    sys.exit(main())
