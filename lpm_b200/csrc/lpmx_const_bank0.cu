// lpmx_const_bank0.cu -- constant bank 0 of the constant-bank velocity path: its own translation unit = its own module = its own
// 64 KB user constant bank (lpmx_const_bank.cuh, lpmx_const_stream.cu).
#define LPMX_CS_BANK 0
#include "lpmx_const_bank.cuh"
