"""Emit C++ raw-string constants holding the given header files (used by jit.cuh for NVRTC)."""
import sys
from pathlib import Path

for name in sys.argv[1:]:
    text = Path(name).read_text()
    assert ')QSVJIT"' not in text
    ident = "kJitSrc_" + Path(name).stem
    # raw string literals are limited to 64 KB per literal on some compilers: split by chunks of lines
    lines = text.splitlines(keepends=True)
    print(f"static const char {ident}[] =")
    for i in range(0, len(lines), 200):
        print('R"QSVJIT(' + "".join(lines[i:i + 200]) + ')QSVJIT"')
    print(";")
