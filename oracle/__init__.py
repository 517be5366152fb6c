"""TEST INFRASTRUCTURE ONLY (parity unpinned, see torch_oracle.py header).

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Never from the product package."""
