// Shim (ours): included by cseq_comparator.cpp, nothing of it is used.
#pragma once
