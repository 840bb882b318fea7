/* Compiled (as plain C11) and linked against libdips_b200.so by tests/test_abi.py: proves that include/dips_b200.h is a
 * valid C header and that a C caller -- i.e. any FFI -- can use the library.  Touches no device: with no GPU
 * dipsb_create must fail cleanly, with a GPU it creates and destroys a context. */
#include <stdio.h>
#include <string.h>

#include "dips_b200.h"

int main(void) {
    dipsb_config cfg;
    dipsb_ctx *ctx = NULL;
    uint32_t plan[8];
    if (dipsb_abi_version() != DIPSB_ABI_VERSION) return 2;
    dipsb_default_config(&cfg);
    if (cfg.struct_size != sizeof cfg) return 3;
    if (dipsb_plan_query(1920, 1080, DIPSB_FMT_RGB8, 148, plan) != DIPSB_OK) return 4;
    printf("plan: %u tiles x %u px, %u threads, %u stages\n", plan[0], plan[5], plan[2], plan[3] & 0xFFFF);
    cfg.width = 64; cfg.height = 48; cfg.format = DIPSB_FMT_RGBX8;
    int32_t rc = dipsb_create(&cfg, &ctx);
    if (rc == DIPSB_OK) {
        printf("context created\n");
        if (dipsb_frames_processed(ctx) != 0) return 5;
        dipsb_destroy(ctx);
    } else {
        const char *msg = dipsb_last_error(NULL);
        printf("create failed as expected without a GPU: %s\n", msg);
        if (rc >= 0 || !msg || !strlen(msg)) return 6;
    }
    return 0;
}
