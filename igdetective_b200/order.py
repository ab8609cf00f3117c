"""Row order of the reference's RSS lists.

find_valid_rss iterates `first_set`, a Python set that was built in one pool worker, pickled to
the parent and pickled again into another worker (py/IGDetective.py:144,162,196-204).  Its
iteration order is a property of CPython's set, neither sorted nor insertion order, and it decides
the row order of rss_*.csv / genes_*.tsv.  To be byte-identical the glue replays exactly those
container operations on the ascending index list the GPU scan returns; nothing here touches
sequence data."""
import pickle


def reference_set_order(ascending):
    s = set(ascending)
    s = pickle.loads(pickle.dumps(s))
    s = pickle.loads(pickle.dumps(s))
    return list(s)


def rank_map(ascending):
    return {v: i for i, v in enumerate(reference_set_order(ascending))}
