"""Helpers shared by the parity tests: load the committed reference fixtures."""
import glob
import gzip
import json
import os

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIELDS = ["r", "avail", "pages", "nab", "ap", "fp", "rb", "shrink", "mapped"]


def golden_names():
    names = sorted(os.path.basename(p)[:-len(".json.gz")] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.json.gz")))
    return [n for n in names if not n.startswith("pool_")]   # pool_*: prefix-cache fixtures (test_pool_*.py)


def load_golden(name):
    with gzip.open(os.path.join(GOLDEN_DIR, name + ".json.gz"), "rb") as f:
        fix = json.loads(f.read().decode())
    fix["records"] = [dict(zip(FIELDS, r)) for r in fix["records"]]
    for r in fix["records"]:
        r["shrink"] = bool(r["shrink"])
    return fix


def diff_record(got: dict, want: dict):
    """Names of the fields that differ (empty list = identical)."""
    return [k for k in FIELDS if got.get(k) != want.get(k)]
