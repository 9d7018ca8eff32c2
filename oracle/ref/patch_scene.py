"""Generates a GCC-13-compilable scene.cpp from the reference's src/scene.cpp at build time.

The reference initialises `Intersection` (a class with a user-declared constructor,
/root/reference/include/intersection.h:27-37) with C99 designated initialisers at
src/scene.cpp:207-219 and :322-334.  Old compilers accepted that as a positional
constructor call; GCC 13 rejects it.  This rewrites exactly those two statements into
the positional constructor call with the same expressions in the same order.  Nothing
else is touched and the output lives only under oracle/_ref/build (git-ignored).
"""
import re
import sys

src = open(sys.argv[1]).read()
pat = re.compile(r"Intersection hit = \{(.*?)\n(\s*)\};", re.S)


def repl(m):
    args = []
    for line in m.group(1).split("\n"):
        s = line.strip()
        if not s or s.startswith("//"):
            continue
        mm = re.match(r"\.(\w+)\s*=\s*(.*?),?$", s)
        assert mm, s
        args.append(mm.group(2))
    assert len(args) == 9, args
    return "Intersection hit(\n" + ",\n".join(m.group(2) + "    " + a for a in args) + "\n" + m.group(2) + ");"


out, n = pat.subn(repl, src)
assert n == 2, n
open(sys.argv[2], "w").write(out)
