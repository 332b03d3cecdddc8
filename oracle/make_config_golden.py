"""Regenerates tests/golden/config_fields.json from the live reference (/root/reference): every UPPER_CASE attribute of the three
DAD config modules (I/config.py, C/config_casia.py, E/config_emodb.py) with its value; paths and non-scalars are recorded by name only.
TEST INFRASTRUCTURE ONLY.  Usage: python oracle/make_config_golden.py"""
import importlib.util
import json
import os
import sys

REF = os.environ.get("DAD_REFERENCE", "/root/reference")
FILES = {"iemocap": "IEMOCAP/DAD-train-IEMOCAP/config.py", "casia": "CASIA/DAD-train-CASIA/config_casia.py",
         "emodb": "EMODB/DAD-train-EMODB/config_emodb.py"}


def main():
    sys.dont_write_bytecode = True
    out = {}
    for tag, rel in FILES.items():
        spec = importlib.util.spec_from_file_location("refcfg_" + tag, os.path.join(REF, rel))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        d = {}
        for k, v in vars(m).items():
            if not k.isupper():
                continue
            scalar = isinstance(v, (int, float, bool, str, list, dict)) and not (isinstance(v, str) and ("/" in v or "\\" in v))
            d[k] = v if scalar else "<non-scalar or path>"
        out[tag] = d
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "config_fields.json")
    json.dump(out, open(dst, "w"), indent=1, sort_keys=True)
    print({k: len(v) for k, v in out.items()}, "->", dst)


if __name__ == "__main__":
    main()
