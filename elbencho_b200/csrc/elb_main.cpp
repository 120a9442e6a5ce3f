/*
 * elbencho-b200 executable: everything lives in libelbencho_b200.so (elb_cli_main), found next to
 * this binary through its $ORIGIN rpath. Reference counterpart: source/Main.cpp:13-68.
 */
#include "elbencho_b200.h"

int main(int argc, char** argv)
{
	return elb_cli_main(argc, argv);
}
