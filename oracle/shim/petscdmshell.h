// TEST INFRASTRUCTURE (oracle): empty stand-in, see petscksp.h
