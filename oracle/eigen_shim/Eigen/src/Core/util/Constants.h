// oracle/_ref build only: /root/reference/src/Clusterer.cpp:18 includes this path; nothing from it is used.
