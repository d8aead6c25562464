/*
 * qs_hostio.h - host-side plumbing of the host entry points (qs_cuda.cu): a small thread pool
 * for the gather / scatter of libjpeg's block rows and single-thread work queues that keep the
 * upload and the download sides of a run going while the calling thread enqueues kernels.
 *
 * Why: libjpeg hands the coefficient arrays over as separately allocated, pageable block rows
 * (access_virt_barray, reference quantsmooth.h:2592-2594).  A DMA engine wants
 * pinned memory, so every byte crosses a host-side memcpy into (and back out of) a pinned staging
 * buffer.  Round 1 did that single-threaded before / after the device run and paid
 * cudaHostAlloc + cudaFreeHost per call; here the staging buffer lives in the context
 * (grow-only) and the copies run band by band on worker threads, overlapped with the H2D / D2H
 * copies and the kernels of the slab pipeline.
 */
#ifndef QS_HOSTIO_H
#define QS_HOSTIO_H

#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

/* fixed pool; parallel_for may be called from several threads at once */
class QsPool {
public:
	explicit QsPool(int nthreads) : stop_(false) {
		if (nthreads < 1) nthreads = 1;
		for (int i = 0; i < nthreads; i++) th_.emplace_back([this] { loop(); });
	}
	~QsPool() {
		{ std::lock_guard<std::mutex> l(m_); stop_ = true; }
		cv_.notify_all();
		for (std::thread &t : th_) t.join();
	}
	int size() const { return (int)th_.size(); }
	/* runs fn(i) for i in [0, n) on the pool threads and the caller; returns when all are done */
	void parallel_for(int n, const std::function<void(int)> &fn) {
		if (n <= 0) return;
		Batch b; b.fn = &fn; b.n = n; b.next = 0; b.done = 0; b.users = 0;
		{
			std::lock_guard<std::mutex> l(m_);
			q_.push_back(&b);
		}
		cv_.notify_all();
		int mine = run_items(&b);                        /* the caller helps */
		std::unique_lock<std::mutex> l(m_);
		b.done += mine;
		/* the batch lives on this stack frame: leave only when no pool thread holds it any more */
		b.cv.wait(l, [&] { return b.done == b.n && b.users == 0; });
		for (auto it = q_.begin(); it != q_.end(); ++it) if (*it == &b) { q_.erase(it); break; }
	}
private:
	struct Batch {
		const std::function<void(int)> *fn; int n;
		std::atomic<int> next;
		int done, users;                                 /* guarded by m_ */
		std::condition_variable cv;
	};
	static int run_items(Batch *b) {
		int mine = 0;
		for (;;) {
			int i = b->next.fetch_add(1);
			if (i >= b->n) break;
			(*b->fn)(i); mine++;
		}
		return mine;
	}
	void loop() {
		std::unique_lock<std::mutex> l(m_);
		for (;;) {
			Batch *b = NULL;
			cv_.wait(l, [&] {
				if (stop_) return true;
				for (Batch *x : q_) if (x->next.load() < x->n) { b = x; return true; }
				return false;
			});
			if (stop_) return;
			if (!b) continue;
			b->users++;
			l.unlock();
			int mine = run_items(b);
			l.lock();
			b->users--; b->done += mine;
			if (b->done == b->n && b->users == 0) b->cv.notify_all();
		}
	}
	std::mutex m_; std::condition_variable cv_;
	std::deque<Batch*> q_; bool stop_;
	std::vector<std::thread> th_;
};

/* one thread, FIFO of jobs; drain() waits until the queue is empty and the thread idle */
class QsWorker {
public:
	QsWorker() : stop_(false), busy_(false), th_([this] { loop(); }) {}
	~QsWorker() {
		{ std::lock_guard<std::mutex> l(m_); stop_ = true; }
		cv_.notify_all();
		th_.join();
	}
	void post(std::function<void()> f) {
		{ std::lock_guard<std::mutex> l(m_); q_.push_back(std::move(f)); }
		cv_.notify_all();
	}
	void drain() {
		std::unique_lock<std::mutex> l(m_);
		idle_.wait(l, [&] { return q_.empty() && !busy_; });
	}
private:
	void loop() {
		for (;;) {
			std::function<void()> f;
			{
				std::unique_lock<std::mutex> l(m_);
				cv_.wait(l, [&] { return stop_ || !q_.empty(); });
				if (q_.empty()) return;                  /* stop requested and nothing left */
				f = std::move(q_.front()); q_.pop_front(); busy_ = true;
			}
			f();
			{
				std::lock_guard<std::mutex> l(m_);
				busy_ = false;
			}
			idle_.notify_all();
		}
	}
	std::mutex m_; std::condition_variable cv_, idle_;
	std::deque<std::function<void()> > q_; bool stop_, busy_;
	std::thread th_;
};

#endif
