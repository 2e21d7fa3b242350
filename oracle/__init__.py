"""ORACLE — CPU restatement of the chanshing/cfd hot path.  Test infrastructure only (see oracle.cpp)."""
