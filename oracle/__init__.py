"""ORACLE — test infrastructure only.  See oracle/flat_ip.py for scope and parity status."""
