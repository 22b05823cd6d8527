#!/usr/bin/env python
"""Regenerates tests/golden/golden.json and tests/golden/*.zl from the UNMODIFIED reference (oracle/_ref, built
by oracle/Makefile from /root/reference).  Run in the container that has /root/reference:
    python tests/golden/make_golden.py
golden.json: for every named input of tests/_inputs.py and every level e0-e4: input md5/size, compressed
md5/size, and the (encpos, rlen, olen) triple of every sub-block.  *.zl: a few small compressed streams kept
verbatim as decoder known-answer inputs."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from _inputs import small_cases, block_boundary_cases  # noqa: E402
from _libs import Ref, walk_container  # noqa: E402

KEEP = {"one", "three", "hello", "a1000", "bytes256x64", "len277", "text64k", "period3"}


def main():
    ref = Ref()
    out = {}
    for name, data in small_cases() + block_boundary_cases():
        rec = {"size": len(data), "md5": hashlib.md5(data).hexdigest(), "levels": {}}
        for level in range(5):
            z = ref.encode(data, level)
            rec["levels"][str(level)] = {
                "size": len(z), "md5": hashlib.md5(z).hexdigest(),
                "subblocks": [[b, e, r, o] for (b, e, r, o, _) in walk_container(z)][:64],
            }
            if name in KEEP and level in (0, 4):
                with open(os.path.join(HERE, "%s.e%d.zl" % (name, level)), "wb") as f:
                    f.write(z)
        out[name] = rec
        print(name, rec["size"], [rec["levels"][str(l)]["size"] for l in range(5)])
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
