"""format constants: rule-generated tables == the reference's generated .inc files (src/tables/*.inc)"""
import os
import re

import numpy as np
import pytest

REF_TABLES = "/root/reference/src/tables"


def _inc(name):
    txt = open(os.path.join(REF_TABLES, name)).read()
    txt = re.sub(r"//[^\n]*", "", txt)
    return np.array([int(x) for x in re.findall(r"\d+", txt)])


@pytest.mark.skipif(not os.path.isdir(REF_TABLES), reason="reference tree not on this machine")
def test_oracle_tables_match_reference_inc(oracle):
    assert np.array_equal(oracle.table("mtfinit", 256), _inc("table_mtfinit.inc"))
    assert np.array_equal(oracle.table("mtfnext", 256), _inc("table_mtfnext.inc"))
    assert np.array_equal(oracle.table("idx_code", 4096), _inc("table_matchidx_code.inc"))
    assert np.array_equal(oracle.table("idx_base", 32), _inc("table_matchidx_base.inc"))
    assert np.array_equal(oracle.table("idx_bits", 32), _inc("table_matchidx_blen.inc"))


def test_tables_self_consistent(oracle):
    init = oracle.table("mtfinit", 256)
    assert sorted(init.tolist()) == list(range(256))
    code, base, bits = oracle.table("idx_code", 4096), oracle.table("idx_base", 32), oracle.table("idx_bits", 32)
    for i in range(4096):
        c = code[i]
        assert base[c] <= i < base[c] + (1 << bits[c])
    nxt = oracle.table("mtfnext", 256)
    assert all(nxt[i] <= i for i in range(256)) and nxt[0] == 0 and nxt[255] == 140
