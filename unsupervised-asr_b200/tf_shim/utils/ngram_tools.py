"""`ngram2kernel` / `read_ngram` / `load_vocab` with the reference's signatures (utils/tools.py:255-279, 365-374;
utils/dataProcess.py:6-17) for a TF process: `from utils.tools import ngram2kernel` can be pointed here.  The functions
live in eodm_b200/tools.py, which needs numpy only -- importing them does not import PyTorch (tests/test_host_cpu.py)."""
from eodm_b200.tools import load_vocab, ngram2kernel, read_ngram  # noqa: F401
