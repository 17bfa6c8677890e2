"""TEST INFRASTRUCTURE ONLY.  CPU oracle for the VideoSwap denoising hot path.  Nothing under
videoswap_b200/ may import this package; only tests/, __graft_entry__.smoke() and bench.py's CPU legs do."""
