#!/usr/bin/env python3
"""Generates the three headers the reference's cmake step would configure (oracle build only).

  CMakeConfig.h   <- CMakeConfig.h.in        (/root/reference/CMakeLists.txt:60-67)
  nlopt_config.h  <- nlopt/nlopt_config.h.in (values for x86-64 glibc)
  nlopt.hpp       <- nlopt/src/api/nlopt-in.hpp with the enum block expanded from nlopt.h
                     (the GEN_ENUMS_HERE rule of nlopt/src/api/CMakeLists.txt:66-79)
"""
import re
import sys

ref, out = sys.argv[1], sys.argv[2]

with open(f"{out}/CMakeConfig.h", "w") as f:
    f.write("#ifndef CMAKECONFIG_H\n#define CMAKECONFIG_H\n"
            "#define RESEQ_VERSION_MAJOR 1\n#define RESEQ_VERSION_MINOR 1\n"
            f"#define PROJECT_SOURCE_DIR \"{ref}\"\n#endif\n")

with open(f"{out}/nlopt_config.h", "w") as f:
    f.write("""#ifndef NLOPT_CONFIG_H
#define NLOPT_CONFIG_H
#define BUGFIX_VERSION 0
#define MAJOR_VERSION 2
#define MINOR_VERSION 5
#define HAVE_COPYSIGN
#define HAVE_FPCLASSIFY
#define HAVE_GETOPT_H
#define HAVE_GETPID
#define HAVE_GETTIMEOFDAY
#define HAVE_INTTYPES_H
#define HAVE_ISINF
#define HAVE_ISNAN
#define HAVE_QSORT_R
#define HAVE_STDINT_H
#define HAVE_STDLIB_H
#define HAVE_STRINGS_H
#define HAVE_STRING_H
#define HAVE_SYS_STAT_H
#define HAVE_SYS_TYPES_H
#define HAVE_SYS_TIME_H
#define HAVE_TIME
#define HAVE_UINT32_T
#define HAVE_UNISTD_H
#define SIZEOF_UNSIGNED_INT 4
#define SIZEOF_UNSIGNED_LONG 8
#define THREADLOCAL __thread
#define TIME_WITH_SYS_TIME 1
#endif
""")

enum_lines = [l.rstrip("\n") for l in open(f"{ref}/nlopt/src/api/nlopt.h") if re.search(r"    NLOPT_[A-Z0-9_]+", l)]
with open(f"{out}/nlopt.hpp", "w") as f:
    for line in open(f"{ref}/nlopt/src/api/nlopt-in.hpp"):
        f.write(line)
        if "GEN_ENUMS_HERE" in line:
            f.write("  enum algorithm {\n")
            for l in enum_lines:
                f.write(l.replace("NLOPT_", "") + "\n")
                if "NLOPT_NUM_ALGORITHMS" in l:
                    f.write("  };\n  enum result {\n")
                elif "NLOPT_NUM_RESULTS" in l:
                    f.write("  };\n")
