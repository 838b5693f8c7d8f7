// stands in for include/Tracking.h (Pangolin; src/pnpmatch.cc includes it and uses nothing of it)
