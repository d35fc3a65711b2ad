"""TEST INFRASTRUCTURE ONLY.  Nothing in rpo_b200/ may import this package (see oracle/README.md)."""
