// device_b200.cpp -- see device_b200.h
#include "device_b200.h"

#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <string>
#include <thread>

namespace mc2i {

namespace {

int device_index()
{
	const char *dev_env = std::getenv("MC2_DEVICE");
	return dev_env ? std::atoi(dev_env) : 0;
}

struct EarlyContext {
	std::thread th;
	std::mutex mu;
	mc2_ctx *ctx = nullptr;
	std::string err;
	EarlyContext()
	{
		if (std::getenv("MC2_NO_PREWARM")) {
			return;
		}
		th = std::thread([this]() { create(); });
	}
	void create()
	{
		if (mc2_ctx_create(device_index(), &ctx) != MC2_OK) {
			err = mc2_last_error();
			ctx = nullptr;
		}
	}
	mc2_ctx *get()
	{
		std::lock_guard<std::mutex> lock(mu);
		if (th.joinable()) {
			th.join();
		}
		if (!ctx && err.empty()) {
			create();
		}
		if (!ctx) {
			std::cerr << "meshclust2_b200: " << err << std::endl;
			throw std::runtime_error(err);
		}
		return ctx;
	}
	~EarlyContext()
	{
		if (th.joinable()) {
			th.join();
		}
		// The context is NOT destroyed here: static objects of other translation units (the Trainer's prewarm threads,
		// the device mirrors) may still be using it, and their destruction order against this one is unspecified.  The
		// process is exiting; the driver releases the context.
	}
} g_early;

std::mutex g_device_mu;

} // namespace

mc2_ctx *shared_ctx()
{
	return g_early.get();
}

std::mutex &device_mutex()
{
	return g_device_mu;
}

bool batching_enabled()
{
	static const bool off = std::getenv("MC2_NO_BATCH") != nullptr;
	return !off;
}

void ok(int rc)
{
	if (rc != MC2_OK) {
		std::cerr << "meshclust2_b200: " << mc2_last_error() << std::endl;
		throw std::runtime_error(mc2_last_error());
	}
}

} // namespace mc2i
