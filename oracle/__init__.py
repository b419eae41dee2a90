"""CPU oracle of the TSP-GNN hot path.  TEST INFRASTRUCTURE ONLY (see tspgnn_oracle.py):
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs,
never by the product package."""
