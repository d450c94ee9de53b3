"""The C-ABI header is plain C: a C translation unit that includes it and references every entry
point compiles with gcc -std=c99 -Wall -Werror and links against the library."""
import os
import re
import subprocess

from conftest import ROOT


def test_header_compiles_as_c_and_links(tmp_path):
    from pylogeny_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "pylogeny_b200.h")).read()
    names = sorted(set(re.findall(r"\b(plg_[a-z0-9_]+)\s*\(", re.sub(r"/\*.*?\*/", "", hdr, flags=re.S))))
    src = tmp_path / "use.c"
    body = "\n".join("  p[%d] = (void *)%s;" % (i, n) for i, n in enumerate(names))
    src.write_text('#include "pylogeny_b200.h"\n#include <stdio.h>\nint main(void) {\n  void *p[%d];\n%s\n'
                   '  printf("%%d %%p\\n", plg_version(), p[0]);\n  return 0;\n}\n' % (len(names), body))
    exe = tmp_path / "use"
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-Wno-pedantic", "-I", os.path.join(ROOT, "include"),
           str(src), "-o", str(exe), "-L", libdir, "-l:libpylogeny_b200.so", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    out = subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True)
    assert out.returncode == 0 and out.stdout.split()[0] == "100"
