"""include/mcb200.h is a plain-C header: it must compile as C99 and as C++ without CUDA headers."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "use.c"
    src.write_text('#include "mcb200.h"\n'
                   'int main(void) { mcb_option_data o = {100, 1, 100, 0.05f, 0.2f, 120, 10, 50, 1000, 10, 1, 1.0f};\n'
                   '  mcb_result r; mcb_engine *e = 0; (void)o; (void)r; (void)e;\n'
                   '  return sizeof(mcb_option_data) == 48 && sizeof(mcb_result) == 40 ? 0 : 1; }\n')
    for compiler, std in (("gcc", "-std=c99"), ("g++", "-std=c++11")):
        exe = tmp_path / ("use_" + compiler)
        cmd = [compiler, std, "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include")]
        if compiler == "g++":
            cmd += ["-x", "c++"]
        subprocess.run(cmd + [str(src), "-o", str(exe)], check=True)
        assert subprocess.run([str(exe)]).returncode == 0
