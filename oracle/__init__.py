"""ORACLE — test infrastructure only (see oracle/README.md).  Nothing under marqo_b200/ may import this package."""
