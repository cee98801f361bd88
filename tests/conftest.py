import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GTREF = os.path.join(ROOT, "oracle", "_ref", "gtref")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz"))


class Golden:
    """One committed fixture: raw index files + expected text per minlength."""

    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.minlengths = [int(m) for m in self.z["minlengths"]]
        self.flags = str(self.z["flags"])

    def materialise(self, directory):
        """Write <dir>/<name>.{prj,esq,suf,lcp,llv,bwt} byte for byte."""
        base = os.path.join(str(directory), self.name)
        for sfx in ("prj", "esq", "suf", "lcp", "llv", "bwt"):
            with open(base + "." + sfx, "wb") as fh:
                fh.write(self.z["file_" + sfx].tobytes())
        return base

    def expected(self, minlength, policy="gt"):
        return self.z["exp_%s_%d" % (policy, minlength)].tobytes()

    def expected_rel(self, minlength):
        """The reference's own '<len> <count> <seqnum> <relpos>...' lines (gtref ... rel:
        gt_encseq_seqnum / gt_encseq_seqstartpos), gt policy."""
        return self.z["exp_rel_%d" % minlength].tobytes()

    def tables(self):
        from oracle import smax_oracle as O
        prj = {}
        for line in self.z["file_prj"].tobytes().decode().splitlines():
            if line.startswith("dbfile=") or "=" not in line:
                continue
            k, v = line.split("=", 1)
            prj[k] = float(v) if k == "averagelcp" else int(v)
        n = prj["numberofallsortedsuffixes"]
        lcp = self.z["file_lcp"]
        bwt = self.z["file_bwt"]
        sufraw = self.z["file_suf"]
        suf = np.frombuffer(sufraw.tobytes(), dtype=np.uint64 if sufraw.size == 8 * n else np.uint32)
        llv = np.frombuffer(self.z["file_llv"].tobytes(), dtype=O.LLV_DTYPE)
        return O.EsaTables(prj, lcp, bwt, llv, suf)


@pytest.fixture(scope="session")
def c_oracle():
    from oracle import smax_oracle as O
    O.build_c_oracle()
    return O


@pytest.fixture(scope="session")
def libsmax():
    from genometools_smax_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        from genometools_smax_b200 import _build
        _build.build()
    return capi
