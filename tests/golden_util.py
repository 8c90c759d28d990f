import gzip
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with gzip.open(os.path.join(GOLDEN, name + ".json.gz"), "rb") as fh:
        return json.loads(fh.read().decode("ascii"))
