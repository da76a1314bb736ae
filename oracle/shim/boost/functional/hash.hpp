// oracle/shim/boost/functional/hash.hpp -- kino_astar.h includes it but hashes with std::hash (kino_astar.h:61-73)
#pragma once
