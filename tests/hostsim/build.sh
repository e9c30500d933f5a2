#!/bin/sh
# Builds the host test double (see hostsim.cpp).  Output stays under tests/hostsim/.
set -e
cd "$(dirname "$0")"
g++ -O2 -std=c++17 -ffp-contract=off -shared -fPIC -o libhostsim.so hostsim.cpp
