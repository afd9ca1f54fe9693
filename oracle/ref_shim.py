#!/usr/bin/env python3
"""Launch the UNMODIFIED reference indexer.py / merger.py from /root/reference.

TEST INFRASTRUCTURE ONLY, and build-container only: /root/reference does not
exist on the GPU box, so nothing under tests -m gpu, smoke() or bench.py may
call this.  It is used by oracle/make_golden.py to manufacture tests/golden/.

Nothing arithmetic is touched.  The shims:
  1. `import bgzip` (tools.py:17) -> empty module; the name is never used.
  2. Header.__init__ accepts the `sample_name=` keyword that HEAD indexer.py
     passes (indexer.py:311-322) but HEAD tools.py does not take (tools.py:111).
  3. JSON encoder casts numpy integers (np.count_nonzero returns np.int64 under
     NumPy 2, tools.py:256,261).
  4. merger only: time.sleep(10) in the polling loop (merger.py:181) is
     shortened so a golden run does not idle for 10-20 s.
Run with CWD = the reference directory (tools.py:285 hashes "tools.py" in CWD).

usage: ref_shim.py indexer <fasta> <sample> <K>
       ref_shim.py merger  <project> <kin> <kin> ... [merger flags]
"""
import json
import os
import sys
import types

REF = os.environ.get("PYKMER_REFERENCE", "/root/reference")


def main() -> None:
    import numpy as np

    tool, args = sys.argv[1], sys.argv[2:]
    sys.path.insert(0, REF)
    os.chdir(REF)
    sys.modules.setdefault("bgzip", types.ModuleType("bgzip"))                      # shim 1
    import tools

    init = tools.Header.__init__

    def patched(self, project_name, *a, sample_name=None, **kw):                    # shim 2
        self.sample_name = sample_name
        init(self, project_name, *a, **kw)

    tools.Header.__init__ = patched
    if tool == "indexer":
        default = json.JSONEncoder.default                                           # shim 3
        json.JSONEncoder.default = \
            lambda s, o: int(o) if isinstance(o, np.integer) else default(s, o)
        import indexer
        sys.argv = ["indexer.py"] + args
        indexer.main()
    elif tool == "merger":
        import time
        real_sleep = time.sleep
        sys.argv = ["merger.py"] + args
        import merger
        merger.time.sleep = lambda s: real_sleep(0.05)                               # shim 4
        merger.main()
    else:
        raise SystemExit(__doc__)


if __name__ == "__main__":
    main()
