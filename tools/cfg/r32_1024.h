#define OFFTB_CFG_1024 32, 32, 32, 1, 1, 1, 0, 0, 1, 0, 0
#define OFFTB_MAXT_1024 256
#define OFFTB_MINB_1024 1
