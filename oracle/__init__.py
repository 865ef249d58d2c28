"""CPU oracle for the fauxgl DrawMesh path.  TEST INFRASTRUCTURE ONLY: imported
by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs -- never by the product package."""
