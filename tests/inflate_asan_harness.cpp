// Test harness (tests/test_bam_io.py): runs csrc/smc_inflate.h over a file of valid and corrupted DEFLATE streams with exactly-sized
// heap buffers, built with -fsanitize=address,undefined, so that any read or write outside the documented bounds aborts.
#include "smc_inflate.h"
#include <cstdio>
#include <cstdlib>
#include <vector>
// cases file: repeated [u32 in_len][u32 out_len][u8 expect_ok][in bytes][out bytes if expect_ok]
int main(int argc, char** argv) {
    FILE* f = fopen(argv[1], "rb");
    uint32_t il, ol; uint8_t ok;
    long n = 0, good = 0, rej = 0, wrong = 0;
    while (fread(&il, 4, 1, f) == 1) {
        if (fread(&ol, 4, 1, f) != 1 || fread(&ok, 1, 1, f) != 1) return 2;
        uint8_t* in = (uint8_t*)malloc(il + 64);          // exactly the documented slack
        memset(in + il, 0x5A, 64);
        if (il && fread(in, 1, il, f) != il) return 2;
        std::vector<uint8_t> want(ol);
        if (ok && ol && fread(want.data(), 1, ol, f) != ol) return 2;
        uint8_t* out = (uint8_t*)malloc(ol ? ol : 1);     // exactly out_len: ASan sees any overrun
        int rc = smc_inflate_raw(in, il, out, ol);
        ++n;
        if (rc == 0) { ++good; if (ok && ol && memcmp(out, want.data(), ol) != 0) { ++wrong; } }
        else ++rej;
        if (ok && rc != 0) { printf("case %ld: valid stream rejected\n", n); return 3; }
        free(in); free(out);
    }
    printf("%ld cases: %ld decoded, %ld rejected, %ld wrong\n", n, good, rej, wrong);
    return wrong ? 4 : 0;
}
